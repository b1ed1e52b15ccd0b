"""The adversarial train-step loop bodies of the reference's ``utils/trainer.py``
as callable steps on the libpcadv-backed modules.

The reference's ``run_training_seg`` (utils/trainer.py:849-1026) can drive the
modules unchanged; these functions are the same loop body (:873-966) without
the logging / checkpoint / evaluation code around it, so a benchmark or a
data-parallel launcher can call one iteration.  Losses come back as device
tensors: the caller decides when to synchronise (the reference calls ``.item()``
four times per iteration, :900, :925, :948, :963).
"""
import os

import torch
import torch.nn.functional as F

from .models._chain import weight_cache
from .utils.utils import make_D_label
from .utils.image_pool import ImagePool


def _device_label(like, value, random):
    """Device-side variant of make_D_label: same
    distributions, drawn by the device generator instead of the CPU one."""
    if not random:
        return torch.full_like(like, float(value))
    lo, hi = (0.0, 0.305) if value == 0 else (0.7, 1.05)
    return torch.empty_like(like).uniform_(lo, hi)


def adversarial_seg_step(*a, **kw):
    """One iteration of run_training_seg (utils/trainer.py:873-966); see ``_adversarial_seg_step``.
    The engine copies of the weights are shared by the passes of the step (``weight_cache``)."""
    with weight_cache():
        return _adversarial_seg_step(*a, **kw)


def adversarial_seg_step_fused(*a, **kw):
    """The same iteration through the fused loss heads; see ``_adversarial_seg_step_fused``."""
    with weight_cache():
        return _adversarial_seg_step_fused(*a, **kw)


def _adversarial_seg_step(model, model_D, gan_loss, seg_loss, optimizer, optimizer_D, batch_gt,
                          batch_nogt, args, history_pool_gt=None, history_pool_nogt=None,
                          device_labels=False, label_fn=None):
    """One iteration of run_training_seg (utils/trainer.py:873-966).

    batch_gt = (pts B x N x 3, cls B x 1 x 16, seg B x N), batch_nogt = (pts, cls),
    already on ``args.device``.  ``args`` needs ``device``, ``lambda_seg``,
    ``lambda_adv``.  Returns (l_seg, l_adv, l_D) as 0-d device tensors."""
    gt_label, nogt_label = 1, 0
    pool_gt = history_pool_gt or ImagePool(0)
    pool_nogt = history_pool_nogt or ImagePool(0)

    def label(d_out, value, random):
        if label_fn is not None:
            return label_fn(d_out, value, random)
        if device_labels:
            return _device_label(d_out, value, random)
        return make_D_label(input=d_out, value=value, device=args.device, random=random)

    model.train()
    model_D.train()
    optimizer.zero_grad()
    optimizer_D.zero_grad()

    # ---- train G (:884-929): D frozen
    for param in model_D.parameters():
        param.requires_grad = False
    pts, cls, seg = batch_gt
    pred, _ = model(pts, cls)
    l_seg = seg_loss(pred, seg)
    pred_gt_softmax = F.softmax(pred, dim=1)
    pts_nogt, cls_nogt = batch_nogt
    pred_nogt, _ = model(pts_nogt, cls_nogt)
    pred_nogt_softmax = F.log_softmax(pred_nogt, dim=1)
    D_out = model_D(pred_nogt_softmax)
    loss_adv = gan_loss(D_out, label(D_out, gt_label, False))
    (args.lambda_seg * l_seg + args.lambda_adv * loss_adv).backward()
    _start_reduce(optimizer)

    # ---- train D (:931-963)
    for param in model_D.parameters():
        param.requires_grad = True
    D_out = model_D(pool_gt.query(pred_gt_softmax.detach()))
    loss_D_gt = gan_loss(D_out, label(D_out, gt_label, True)) * 0.5
    loss_D_gt.backward()
    D_out = model_D(pool_nogt.query(pred_nogt_softmax.detach()))
    loss_D_nogt = gan_loss(D_out, label(D_out, nogt_label, True)) * 0.5
    loss_D_nogt.backward()

    optimizer.step()
    optimizer_D.step()
    return l_seg.detach(), loss_adv.detach(), (loss_D_gt + loss_D_nogt).detach()


def _start_reduce(optimizer):
    """Data-parallel runs (parallel.DistributedOptimizer): launch this optimizer's gradient
    all-reduce now, asynchronously; ``optimizer.step()`` waits for it.  A plain optimizer has no
    such hook and nothing happens."""
    start = getattr(optimizer, "reduce_gradients_async", None)
    if start is not None:
        start()


def _adversarial_seg_step_fused(model, model_D, gan_loss, seg_loss, optimizer, optimizer_D, batch_gt,
                                batch_nogt, args, history_pool_gt=None, history_pool_nogt=None,
                                device_labels=False, label_fn=None, one_pass=None, overlap_d=None):
    """The same iteration (utils/trainer.py:873-966) through the generator's fused loss heads
    (SURVEY.md 8f rank 1): ``CrossEntropyLoss`` + ``softmax`` of the labelled pass and
    ``log_softmax`` of the unlabelled pass are one kernel each over the logits, and the
    discriminator reads their packed 16-bit output directly.  ``seg_loss`` must be
    ``nn.CrossEntropyLoss()`` with default arguments (mean over all points); it is accepted only
    to keep the signature of ``adversarial_seg_step``.  Same losses, same gradients.

    ``one_pass`` (default: when the loss weights allow it): the two generator passes of the
    iteration run as ONE pass over the labelled + unlabelled clouds (``forward_ce_logsoftmax``):
    clouds are independent through the network, so the sums are the same, with half the launches and
    no gradient accumulation between two backward passes.

    ``overlap_d`` (default: inside a CUDA-graph capture): the discriminator phase is issued on a
    second stream, concurrently with the generator's backward."""
    if not isinstance(seg_loss, torch.nn.CrossEntropyLoss) or seg_loss.weight is not None or \
            seg_loss.reduction != "mean" or seg_loss.label_smoothing != 0.0:
        raise ValueError("the fused step implements nn.CrossEntropyLoss() with default arguments")
    if 0 <= seg_loss.ignore_index < model.output_dim:
        # the fused head ignores exactly the labels outside [0, k) (the default ignore_index = -100)
        raise ValueError("the fused step cannot ignore a label inside [0, %d)" % model.output_dim)
    gt_label, nogt_label = 1, 0
    pool_gt = history_pool_gt or ImagePool(0)
    pool_nogt = history_pool_nogt or ImagePool(0)

    def label(d_out, value, random):
        if label_fn is not None:
            return label_fn(d_out, value, random)
        if device_labels:
            return _device_label(d_out, value, random)
        return make_D_label(input=d_out, value=value, device=args.device, random=random)

    model.train()
    model_D.train()
    optimizer.zero_grad()
    optimizer_D.zero_grad()

    for param in model_D.parameters():
        param.requires_grad = False
    pts, cls, seg = batch_gt
    pts_nogt, cls_nogt = batch_nogt
    if one_pass is None:
        # One generator pass over the labelled + unlabelled clouds shares one gradient scale, the CE
        # rows'; that needs a CE term and an adversarial term that is not orders of magnitude larger.
        one_pass = (args.lambda_seg > 0 and args.lambda_adv <= 4.0 * args.lambda_seg
                    and pts.shape[1] == pts_nogt.shape[1] and hasattr(model, "forward_ce_logsoftmax"))
    if one_pass:
        l_seg, pred_gt_softmax, pred_nogt_softmax, _ = model.forward_ce_logsoftmax(
            pts, cls, seg, pts_nogt, cls_nogt)                            # :898-901 and :913-914
    else:
        l_seg, pred_gt_softmax, _ = model.forward_ce(pts, cls, seg)          # :898-901
        pred_nogt_softmax, _ = model.forward_logsoftmax(pts_nogt, cls_nogt)   # :913-914
    D_out = model_D(pred_nogt_softmax)
    loss_adv = gan_loss(D_out, label(D_out, gt_label, False))

    def train_D():                                                           # :931-963
        for param in model_D.parameters():
            param.requires_grad = True
        D_gt = model_D(pool_gt.query(pred_gt_softmax.detach()))
        l_gt = gan_loss(D_gt, label(D_gt, gt_label, True)) * 0.5
        l_gt.backward()
        D_nogt = model_D(pool_nogt.query(pred_nogt_softmax.detach()))
        l_nogt = gan_loss(D_nogt, label(D_nogt, nogt_label, True)) * 0.5
        l_nogt.backward()
        return l_gt, l_nogt

    if overlap_d is None:
        overlap_d = _OVERLAP_D and torch.cuda.is_current_stream_capturing()
    if overlap_d:
        # The discriminator phase needs only the two maps the generator's forward produced and D's own
        # weights -- nothing of the generator's backward.  It is a string of small, latency-bound
        # kernels (2^20 rows x <= 128 channels), the generator's backward a string of HBM-bound ones:
        # issued on a second stream they run side by side (inside a graph capture: a parallel branch).
        # Every tensor the branch reads stays referenced until the join below (the step's locals and
        # the step scope's forward cache), so no block is recycled under it; the smoothed labels are
        # drawn in the reference's order (real, then fake) either way.
        main = torch.cuda.current_stream()
        side = _side_stream(main.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            loss_D_gt, loss_D_nogt = train_D()
        (args.lambda_seg * l_seg + args.lambda_adv * loss_adv).backward()    # :927-929
        _start_reduce(optimizer)
        main.wait_stream(side)
    else:
        (args.lambda_seg * l_seg + args.lambda_adv * loss_adv).backward()    # :927-929
        _start_reduce(optimizer)   # G's gradients are final: their all-reduce runs under the D phase
        loss_D_gt, loss_D_nogt = train_D()

    optimizer.step()
    optimizer_D.step()
    return l_seg.detach(), loss_adv.detach(), (loss_D_gt + loss_D_nogt).detach()


_OVERLAP_D = os.environ.get("PCADV_OVERLAP_D", "1") != "0"
_SIDE_STREAMS = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def pointnet_cls_step(model, cls_loss, optimizer, batch, args):
    """One iteration of run_training_pointnet_cls (utils/trainer.py:236-269): PointNetCls forward,
    ``cls_loss`` + ``lambda_regu`` x feature_transform_regularizer when the model has a feature
    transform, backward, optimizer step.  ``batch`` = (pts B x N x 3, labels B) on ``args.device``;
    ``args`` needs ``lambda_cls`` and ``lambda_regu``.  Returns (l_cls, l_regu | None) as device
    tensors (the reference calls ``.item()`` on both, :256-259)."""
    from .models.pointnet import feature_transform_regularizer
    with weight_cache():
        model.train()
        optimizer.zero_grad()
        pts, labels = batch
        pred, _, high_feat = model(pts)
        l_cls = cls_loss(pred, labels)
        l_regu = feature_transform_regularizer(high_feat) if high_feat is not None else None
        loss = args.lambda_cls * l_cls
        if l_regu is not None:
            loss = loss + args.lambda_regu * l_regu
        loss.backward()
        optimizer.step()
    return l_cls.detach(), (l_regu.detach() if l_regu is not None else None)


def pointnet_densecls_step(model, optimizer, batch, args=None):
    """One train iteration of PointNetDenseCls (models/pointnet.py:320-343 with the fix of SURVEY.md
    8c-2; BASELINE.json configs[1]): log-probabilities B x N x k -> ``F.nll_loss`` over all points
    (what the file's upstream, fxia22/pointnet.pytorch, trains the class with -- the reference itself
    has no caller for it), backward, optimizer step.  ``batch`` = (x B x 3 x N, seg B x N int64).
    Returns the loss as a 0-d device tensor."""
    with weight_cache():
        model.train()
        optimizer.zero_grad()
        x, seg = batch
        out, _ = model(x)
        loss = F.nll_loss(out.reshape(-1, out.shape[-1]), seg.reshape(-1))
        loss.backward()
        optimizer.step()
    return loss.detach()


class GraphedStep:
    """Any loop body of this module captured once into a CUDA graph and replayed (the small BASELINE
    configs are launch-bound when issued eagerly).  ``step_fn(*static_inputs)`` must return a tensor
    or a tuple of 0-d tensors (None entries allowed); optimizers must be ``capturable``.  Warm-up
    iterations are undone (see ``GraphedAdversarialSegStep``)."""

    def __init__(self, step_fn, inputs, models, optimizers, warmup=3, restore_state=True):
        self.static = tuple(t.clone() for t in inputs)
        snap = _snapshot_training_state(tuple(models), tuple(optimizers)) if restore_state else None

        def run():
            out = step_fn(*self.static)
            out = out if isinstance(out, (tuple, list)) else (out,)
            return torch.stack([o.float().reshape(()) for o in out if o is not None])

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if snap is not None:
            _restore_training_state(snap, tuple(models), tuple(optimizers))
        from . import _lib
        before = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = run()
        self.launches_per_step = _lib.launch_count() - before

    def __call__(self, *inputs):
        for dst, src in zip(self.static, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out


def adversarial_cls_step(model, model_D, gan_loss, cls_loss, optimizer, optimizer_D, batch_gt, batch_nogt,
                         args, history_pool_gt=None, history_pool_nogt=None, device_labels=False,
                         label_fn=None):
    """One iteration of run_training (utils/trainer.py:426-559): the classification counterpart of
    ``adversarial_seg_step`` -- G = PointNetCls, D = DeepConvDiscNet on ``log_softmax`` of the class
    logits for BOTH discriminator inputs (:472, :492).  batch_gt = (pts, labels), batch_nogt = pts or
    (pts,).  Returns (l_cls, l_adv, l_D) as 0-d device tensors."""
    pool_gt = history_pool_gt or ImagePool(0)
    pool_nogt = history_pool_nogt or ImagePool(0)

    def label(d_out, value, random):
        if label_fn is not None:
            return label_fn(d_out, value, random)
        if device_labels:
            return _device_label(d_out, value, random)
        return make_D_label(input=d_out, value=value, device=args.device, random=random)

    with weight_cache():
        model.train()
        model_D.train()
        optimizer.zero_grad()
        optimizer_D.zero_grad()
        for param in model_D.parameters():                                   # :455-456
            param.requires_grad = False
        pts, labels = batch_gt
        pred, _, _ = model(pts)                                              # :468
        l_cls = cls_loss(pred, labels)
        pred_gt_ls = F.log_softmax(pred, dim=1)                              # :472
        pts_nogt = batch_nogt[0] if isinstance(batch_nogt, (tuple, list)) else batch_nogt
        pred_nogt, _, _ = model(pts_nogt)                                    # :490
        pred_nogt_ls = F.log_softmax(pred_nogt, dim=1)                       # :492
        D_out = model_D(pred_nogt_ls)                                        # :499
        loss_adv = gan_loss(D_out, label(D_out, 1, False))
        (args.lambda_cls * l_cls + args.lambda_adv * loss_adv).backward()    # :511-521

        for param in model_D.parameters():                                   # :524-525
            param.requires_grad = True
        D_out = model_D(pool_gt.query(pred_gt_ls.detach()))                  # :529-531
        loss_D_gt = gan_loss(D_out, label(D_out, 1, True)) * 0.5
        loss_D_gt.backward()
        D_out = model_D(pool_nogt.query(pred_nogt_ls.detach()))              # :545-547
        loss_D_nogt = gan_loss(D_out, label(D_out, 0, True)) * 0.5
        loss_D_nogt.backward()
        optimizer.step()                                                     # :558-559
        optimizer_D.step()
    return l_cls.detach(), loss_adv.detach(), (loss_D_gt + loss_D_nogt).detach()


def _semi_labels(D_out, pred_nogt, semi_TH):
    """Pseudo-labels of the semi-supervised term (utils/trainer.py:727-739, :1998-2010): the
    generator's own argmax where the discriminator output exceeds ``semi_TH``, 255 (ignored)
    elsewhere.  Built on the device (the reference takes the argmax on the host and, in the
    segmentation loop, indexes that host tensor with a device mask -- :2002 -- which only runs on a
    CPU device).  Returns (labels | None when every entry is ignored, fraction kept); like the
    reference (:733) this reads one number back from the device."""
    mask = (D_out.detach() <= semi_TH).squeeze(1)
    semi_gt = torch.argmax(pred_nogt.detach(), dim=1)
    semi_gt = torch.where(mask, torch.full_like(semi_gt, 255), semi_gt)
    ratio = 1.0 - float(mask.sum().item()) / float(mask.numel())
    return (None if ratio == 0.0 else semi_gt), ratio


def adversarial_cls_semi_step(model, model_D, gan_loss, cls_loss, semi_loss, optimizer, optimizer_D, batch_gt,
                              batch_nogt, args, i_iter, history_pool_gt=None, history_pool_nogt=None,
                              label_fn=None):
    """One iteration of run_training_semi (utils/trainer.py:635-794): ``adversarial_cls_step`` plus,
    once ``i_iter > args.semi_start > 0``, ``semi_loss`` (CrossEntropyLoss(ignore_index=255),
    train_classification.py:201) of the unlabelled logits against their own argmax where
    D_out > args.semi_TH.  Returns (l_cls, l_adv, l_semi | None, l_D)."""
    pool_gt = history_pool_gt or ImagePool(0)
    pool_nogt = history_pool_nogt or ImagePool(0)
    label = label_fn or (lambda d_out, value, random: make_D_label(input=d_out, value=value, device=args.device,
                                                                    random=random))
    with weight_cache():
        model.train(); model_D.train()
        optimizer.zero_grad(); optimizer_D.zero_grad()                       # :644-645
        for param in model_D.parameters():
            param.requires_grad = False
        pts, labels = batch_gt
        pred, _, _ = model(pts)                                              # :681
        l_cls = cls_loss(pred, labels)
        pred_gt_ls = F.log_softmax(pred, dim=1)                              # :685
        pts_nogt = batch_nogt[0] if isinstance(batch_nogt, (tuple, list)) else batch_nogt
        pred_nogt, _, _ = model(pts_nogt)                                    # :702
        pred_nogt_ls = F.log_softmax(pred_nogt, dim=1)                       # :704
        D_out = model_D(pred_nogt_ls)                                        # :711
        l_adv = gan_loss(D_out, label(D_out, 1, False))
        loss = args.lambda_cls * l_cls + args.lambda_adv * l_adv
        l_semi = None
        if args.semi_start > 0 and i_iter > args.semi_start:                 # :727
            semi_gt, _ = _semi_labels(D_out, pred_nogt, args.semi_TH)
            if semi_gt is not None:
                l_semi = semi_loss(pred_nogt, semi_gt)
                loss = loss + args.lambda_semi * l_semi
        loss.backward()                                                      # :760
        _start_reduce(optimizer)
        for param in model_D.parameters():
            param.requires_grad = True
        D_out = model_D(pool_gt.query(pred_gt_ls.detach()))                  # :771-773
        loss_D_gt = gan_loss(D_out, label(D_out, 1, True)) * 0.5
        loss_D_gt.backward()
        D_out = model_D(pool_nogt.query(pred_nogt_ls.detach()))              # :785-787
        loss_D_nogt = gan_loss(D_out, label(D_out, 0, True)) * 0.5
        loss_D_nogt.backward()
        optimizer.step()                                                     # :793-794
        optimizer_D.step()
    return l_cls.detach(), l_adv.detach(), (l_semi.detach() if l_semi is not None else None), \
        (loss_D_gt + loss_D_nogt).detach()


def adversarial_seg_semi_step(model, model_D, gan_loss, seg_loss, semi_loss, optimizer, optimizer_D, batch_gt,
                              batch_nogt, args, i_iter, history_pool_gt=None, history_pool_nogt=None,
                              label_fn=None):
    """One iteration of run_training_seg_semi (utils/trainer.py:1927-2061): the adversarial
    segmentation step with the semi-supervised term (:1998-2010); the generator is stepped right
    after its backward (:2022-2023), the discriminator at the end (:2061).
    Returns (l_seg, l_adv, l_semi | None, l_D)."""
    pool_gt = history_pool_gt or ImagePool(0)
    pool_nogt = history_pool_nogt or ImagePool(0)
    label = label_fn or (lambda d_out, value, random: make_D_label(input=d_out, value=value, device=args.device,
                                                                    random=random))
    with weight_cache():
        model.train(); model_D.train()
        optimizer.zero_grad(); optimizer_D.zero_grad()                       # :1937-1938
        for param in model_D.parameters():
            param.requires_grad = False
        pts, cls, seg = batch_gt
        pred, _ = model(pts, cls)                                            # :1969
        l_seg = seg_loss(pred, seg)
        pred_gt_softmax = F.softmax(pred, dim=1)                             # :1972
        pts_nogt, cls_nogt = batch_nogt
        pred_nogt, _ = model(pts_nogt, cls_nogt)                             # :1984
        pred_nogt_softmax = F.log_softmax(pred_nogt, dim=1)                  # :1985
        D_out = model_D(pred_nogt_softmax)                                   # :1987
        l_adv = gan_loss(D_out, label(D_out, 1, False))
        loss = args.lambda_seg * l_seg + args.lambda_adv * l_adv
        l_semi = None
        if args.semi_start > 0 and i_iter > args.semi_start:                 # :1999
            semi_gt, _ = _semi_labels(D_out, pred_nogt, args.semi_TH)
            if semi_gt is not None:
                l_semi = semi_loss(pred_nogt, semi_gt)
                loss = loss + args.lambda_semi * l_semi
        loss.backward()
        optimizer.step()                                                     # :2022-2023
        for param in model_D.parameters():
            param.requires_grad = True
        D_out = model_D(pool_gt.query(pred_gt_softmax.detach()))             # :2030-2032
        loss_D_gt = gan_loss(D_out, label(D_out, 1, True)) * 0.5
        loss_D_gt.backward()
        D_out = model_D(pool_nogt.query(pred_nogt_softmax.detach()))         # :2045-2047
        loss_D_nogt = gan_loss(D_out, label(D_out, 0, True)) * 0.5
        loss_D_nogt.backward()
        optimizer_D.step()                                                   # :2061
    return l_seg.detach(), l_adv.detach(), (l_semi.detach() if l_semi is not None else None), \
        (loss_D_gt + loss_D_nogt).detach()


def adversarial_seg_dual_step(model, sharedDisc, shapeDisc, pointDisc, gan_point_loss, gan_shape_loss, seg_loss,
                              optimizer, optimizer_D_shape, optimizer_D_point, batch_gt, batch_nogt, args,
                              history_pool_gt=None, history_pool_nogt=None, label_fn=None):
    """One iteration of run_training_seg_dual (utils/trainer.py:2150-2284): the generator against a
    shared trunk (BaseDiscNet) with a per-point head (PointDiscNet, BCE-with-logits) and a shape
    head (ShapeDiscNet, CrossEntropyLoss against the cloud's category).  The call sequence is the
    reference's, quirks included: ``optimizer_D_shape.zero_grad()`` twice and ``optimizer_D_point``
    never zeroed (:2171-2172), the generator stepped right after its backward (:2224), the point
    optimizer stepped BEFORE the shape pass runs on the updated shared trunk (:2275-2278).
    Returns (l_seg, l_adv, l_D_point, l_D_shape) as 0-d device tensors."""
    pool_gt = history_pool_gt or ImagePool(0)
    pool_nogt = history_pool_nogt or ImagePool(0)
    label = label_fn or (lambda d_out, value, random: make_D_label(input=d_out, value=value, device=args.device,
                                                                    random=random))
    discs = (sharedDisc, shapeDisc, pointDisc)
    model.train()
    for m in discs:
        m.train()
    optimizer.zero_grad()                                                    # :2170
    optimizer_D_shape.zero_grad()                                            # :2171
    optimizer_D_shape.zero_grad()                                            # :2172
    with weight_cache():
        for m in discs:                                                      # :2174-2179
            for param in m.parameters():
                param.requires_grad = False
        pts, cls, seg = batch_gt
        pred, _ = model(pts, cls)                                            # :2191
        l_seg = seg_loss(pred, seg)
        pred_gt_softmax = F.softmax(pred, dim=1)                             # :2194
        pts_nogt, cls_nogt = batch_nogt
        pred_nogt, _ = model(pts_nogt, cls_nogt)                             # :2206
        pred_nogt_softmax = F.log_softmax(pred_nogt, dim=1)                  # :2207
        D_point = pointDisc(sharedDisc(pred_nogt_softmax))                   # :2209-2210
        loss_adv = gan_point_loss(D_point, label(D_point, 1, False))
        (args.lambda_seg * l_seg + args.lambda_adv * loss_adv).backward()    # :2222-2223
        optimizer.step()                                                     # :2224
        for m in discs:                                                      # :2229-2234
            for param in m.parameters():
                param.requires_grad = True
        D_point = pointDisc(sharedDisc(pool_gt.query(pred_gt_softmax.detach())))       # :2237-2241
        loss_D_point_gt = 0.5 * gan_point_loss(D_point, label(D_point, 1, True))
        D_point = pointDisc(sharedDisc(pool_nogt.query(pred_nogt_softmax.detach())))   # :2253-2257
        loss_D_point_nogt = 0.5 * gan_point_loss(D_point, label(D_point, 0, True))
        (loss_D_point_gt + loss_D_point_nogt).backward()                     # :2273-2274
    optimizer_D_point.step()                                                 # :2275
    with weight_cache():                                                     # sharedDisc just changed
        D_shape = shapeDisc(sharedDisc(pred_gt_softmax.detach()))            # :2277-2278
        cls_gt = cls.argmax(dim=2).squeeze(1)                                # :2279
        loss_D_shape = gan_shape_loss(D_shape, cls_gt.long())                # :2280
        (args.lambda_disc_shape * loss_D_shape).backward()                   # :2283-2284
    optimizer_D_shape.step()                                                 # :2285
    return l_seg.detach(), loss_adv.detach(), (loss_D_point_gt + loss_D_point_nogt).detach(), \
        loss_D_shape.detach()


def _optimizers_of(opt):
    return getattr(opt, "optimizer", opt)


def _snapshot_training_state(models, optimizers):
    """Clones of every parameter and optimizer-state tensor (and the device generator state)."""
    params = [[p.detach().clone() for p in m.parameters()] for m in models]
    states = []
    for opt in optimizers:
        st = _optimizers_of(opt).state
        states.append({id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in d.items()}
                       for p, d in st.items()})
    return params, states, torch.cuda.get_rng_state()


def _restore_training_state(snap, models, optimizers):
    """In-place restore (the tensors keep their addresses, so a later graph capture sees them).
    Optimizer-state tensors that did not exist at snapshot time are zeroed, which is the state a
    fresh Adam / momentum-free SGD starts from."""
    params, states, rng = snap
    with torch.no_grad():
        for m, saved in zip(models, params):
            for p, s in zip(m.parameters(), saved):
                p.copy_(s)
            for p in m.parameters():
                p.grad = None
        for opt, saved in zip(optimizers, states):
            for p, d in _optimizers_of(opt).state.items():
                old = saved.get(id(p), {})
                for k, v in d.items():
                    if torch.is_tensor(v):
                        v.copy_(old[k]) if k in old else v.zero_()
    torch.cuda.set_rng_state(rng)


class GraphedAdversarialSegStep:
    """``adversarial_seg_step`` captured once into a CUDA graph and replayed.

    One iteration launches ~400 kernels (123 libpcadv + the trainer-side torch ops); at cfg3
    sizes the host cannot issue them as fast as the GPU retires them.  The captured graph
    replays the identical kernel sequence from static buffers, so a step costs one launch.
    The smoothed GAN labels keep the reference's semantics (utils/utils.py:22-31: drawn on the
    CPU from torch's default generator): they are drawn into pinned host buffers and copied to
    static device buffers before every replay.  Optimizers must be built with
    ``capturable=True``.

    The constructor runs ``warmup`` iterations on the first batch before capturing (allocator
    warm-up, lazy optimizer state).  With ``restore_state`` (default) the parameters, the optimizer
    state and the device generator are restored in place afterwards, so constructing the object
    does not advance training.  The smoothed labels are drawn by a worker thread from a private generator
    cloned from torch's CPU generator at construction (the reference's sequence of draws; the global
    generator itself is left alone); ``close()`` joins the worker.
    """

    def __init__(self, model, model_D, gan_loss, seg_loss, optimizer, optimizer_D, args, batch_gt,
                 batch_nogt, warmup=3, device_labels=False, fused=False, restore_state=True,
                 history_pool_gt=None, history_pool_nogt=None, one_pass=None, label_draw=None, overlap_d=None,
                 threaded_labels=True):
        for pool in (history_pool_gt, history_pool_nogt):
            if pool is not None and getattr(pool, "pool_size", 0) > 0:
                # the pool's swap decisions are host-side ``random`` draws per sample
                # (utils/image_pool.py:38-52): they cannot be replayed from a captured graph
                raise ValueError("GraphedAdversarialSegStep supports pool_size == 0 only; use "
                                 "adversarial_seg_step(_fused) for history pools")
        self.static_gt = tuple(t.clone() for t in batch_gt)
        self.static_nogt = tuple(t.clone() for t in batch_nogt)
        self.device_labels = device_labels
        # label_draw(real, fake): fills the two pinned host label buffers of the next iteration;
        # default = the reference's draws (utils/utils.py:22-31).  Data-parallel parity tests pass
        # their shard of globally drawn labels here.
        self.label_draw = label_draw
        B, N = batch_nogt[0].shape[0], batch_nogt[0].shape[1]
        Bg = batch_gt[0].shape[0]
        dev = batch_gt[0].device
        self.label_real = torch.empty((Bg, N), dtype=torch.float32, device=dev)
        self.label_fake = torch.empty((B, N), dtype=torch.float32, device=dev)
        # two pinned host slots: the labels of iteration i+1 are drawn while the GPU runs i
        self.host_real = [torch.empty((Bg, N), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.host_fake = [torch.empty((B, N), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._copied = [None, None]
        self._slot = 0

        def label_fn(d_out, value, random):
            if not random:
                return torch.full_like(d_out, float(value))
            if device_labels:
                return _device_label(d_out, value, True)
            return self.label_real if value == 1 else self.label_fake

        step_fn = adversarial_seg_step_fused if fused else adversarial_seg_step

        extra = {"one_pass": one_pass, "overlap_d": overlap_d} if fused else {}

        def run():
            return step_fn(model, model_D, gan_loss, seg_loss, optimizer, optimizer_D,
                           self.static_gt, self.static_nogt, args, label_fn=label_fn, **extra)

        # The smoothed labels travel like the batch: drawn on the host, copied to device staging on a
        # copy stream while the previous iteration computes, moved into the graph's static buffers
        # with a device-to-device copy at the start of the step (no host -> device copy on the
        # critical path, none queued behind the next batch's copy on the DMA engine).
        if not device_labels:
            self.stage_real, self.stage_fake = torch.empty_like(self.label_real), torch.empty_like(self.label_fake)
            self._label_stream = torch.cuda.Stream()
        self._labels_staged = self._labels_consumed = None
        self.threaded_labels = threaded_labels and not device_labels
        self._label_worker = self._label_future = None
        # The labels come from a private generator that starts in the state of torch's CPU generator
        # at construction: the sequence is the one make_D_label would draw in the reference loop
        # (utils/utils.py:26-28) when nothing else consumes that generator, and a background draw
        # can never interleave with other users of the global generator (DataLoader seeds, user code).
        self._label_gen = torch.Generator()
        self._label_gen.set_state(torch.get_rng_state())
        self._draw_labels(0)
        self._stage_labels()
        self._consume_labels()
        # The warm-up iterations (allocator / lazy optimizer state before capture) are real optimizer
        # steps; they must not count as training: parameters, optimizer state and the device
        # generator are put back afterwards, in place, so the first replay is iteration 1 of the
        # reference loop (utils/trainer.py:873) from the caller's state.
        snap = _snapshot_training_state((model, model_D), (optimizer, optimizer_D)) if restore_state else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if snap is not None:
            _restore_training_state(snap, (model, model_D), (optimizer, optimizer_D))
        from . import _lib
        before = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: CUDA calls of other threads (another step object's label worker) must not
        # invalidate this capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.losses = torch.stack(run())
        self.launches_per_step = _lib.launch_count() - before

    # ---- input pipeline: the next batch's host -> device copy runs on a
    # copy stream under the current step; the step then moves it into the graph's static buffers
    # with a device-to-device copy
    def set_jitter(self, sigma=0.01, clip=0.05, seed=0):
        """Augment every prefetched batch on the device: the point coordinates of
        both batches get ``clip(sigma * N(0, 1), -clip, clip)`` added on the copy stream, right behind
        their host -> device copy -- jitter_point_cloud of dataset/modelNetData.py:80-91, which the
        reference applies per cloud on the host inside the DataLoader workers.  The counter-based
        stream advances with every batch, so a (seed, batch number) pair is reproducible."""
        self._jitter = (float(sigma), float(clip), int(seed))
        self._jitter_offset = 0

    def prefetch(self, batch_gt, batch_nogt):
        """Start copying the NEXT step's (pinned) host batch to device staging buffers."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._stage_gt = tuple(torch.empty_like(t) for t in self.static_gt)
            self._stage_nogt = tuple(torch.empty_like(t) for t in self.static_nogt)
            self._staged = torch.cuda.Event()
            self._consumed = None
        cs = self._copy_stream
        if self._consumed is not None:
            cs.wait_event(self._consumed)                  # staging was read by the previous step
        with torch.cuda.stream(cs):
            for dst, src in zip(self._stage_gt + self._stage_nogt, tuple(batch_gt) + tuple(batch_nogt)):
                dst.copy_(src, non_blocking=True)
            if getattr(self, "_jitter", None) is not None:
                from . import ops
                sigma, clip, seed = self._jitter
                for pts in (self._stage_gt[0], self._stage_nogt[0]):
                    ops.jitter(pts, sigma, clip, seed, offset=self._jitter_offset, out=pts)
                    self._jitter_offset += (pts.numel() + 3) // 4
            self._staged.record(cs)

    def step_prefetched(self):
        """Run one step on the batch handed to the last ``prefetch`` call."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        for dst, src in zip(self.static_gt + self.static_nogt, self._stage_gt + self._stage_nogt):
            dst.copy_(src, non_blocking=True)
        self._consumed = torch.cuda.Event()
        self._consumed.record(cur)
        return self()

    def _draw_labels(self, slot):
        """Same draws, same order as make_D_label(random=True) at utils/trainer.py:940-945 and
        :955-960 (CPU generator), into pinned host slot ``slot``."""
        if self.device_labels:
            return
        if self._copied[slot] is not None:
            self._copied[slot].synchronize()          # its previous upload has left the buffer
        if self.label_draw is not None:
            self.label_draw(self.host_real[slot], self.host_fake[slot])
            return
        self.host_real[slot].uniform_(0.7, 1.05, generator=self._label_gen)
        self.host_fake[slot].uniform_(0.0, 0.305, generator=self._label_gen)

    def _stage_labels(self):
        """Host slot -> device staging, on the label copy stream."""
        if self.device_labels:
            return
        slot, ls = self._slot, self._label_stream
        if self._labels_consumed is not None:
            ls.wait_event(self._labels_consumed)           # the previous labels left the staging buffers
        with torch.cuda.stream(ls):
            self.stage_real.copy_(self.host_real[slot], non_blocking=True)
            self.stage_fake.copy_(self.host_fake[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(ls)
        self._copied[slot] = ev
        self._labels_staged = ev

    def _consume_labels(self):
        """Device staging -> the graph's static label buffers, on the current stream."""
        if self.device_labels:
            return
        cur = torch.cuda.current_stream()
        cur.wait_event(self._labels_staged)
        self.label_real.copy_(self.stage_real, non_blocking=True)
        self.label_fake.copy_(self.stage_fake, non_blocking=True)
        self._labels_consumed = torch.cuda.Event()
        self._labels_consumed.record(cur)

    def __call__(self, batch_gt=None, batch_nogt=None):
        """Copy the batch into the static buffers (skipped when None: reuse), draw labels, replay.
        Returns the static 3-element loss tensor (l_seg, l_adv, l_D)."""
        if batch_gt is not None:
            for dst, src in zip(self.static_gt, batch_gt):
                dst.copy_(src, non_blocking=True)
        if batch_nogt is not None:
            for dst, src in zip(self.static_nogt, batch_nogt):
                dst.copy_(src, non_blocking=True)
        self._join_labels()                            # the draw + staging of this iteration's labels
        self._consume_labels()
        self.graph.replay()
        self._slot ^= 1
        self._next_labels()                            # next iteration's labels, under the GPU's shadow
        return self.losses

    # Drawing 2 x B x N uniform floats from the CPU generator (the reference's semantics) takes about as
    # long as the whole cfg5 step on the host (measured: 14 ms for 2 x 2^20 floats), so it must not sit
    # on the launching thread: a single worker thread draws the next iteration's labels (ATen releases
    # the GIL) and enqueues their copy to the device staging buffers; the launching thread only joins it
    # right before it needs them.  One worker, strictly one draw after the other: the sequence of draws
    # from the generator is the reference's (real, fake, real, fake, ...).
    def _next_labels(self):
        if self.device_labels:
            return
        if not self.threaded_labels:
            self._draw_labels(self._slot)
            self._stage_labels()
            return
        if self._label_worker is None:
            from concurrent.futures import ThreadPoolExecutor
            self._label_worker = ThreadPoolExecutor(max_workers=1, thread_name_prefix="pcadv-labels")
        dev = self.label_real.device

        def job():
            torch.cuda.set_device(dev)
            self._draw_labels(self._slot)
            self._stage_labels()
        self._label_future = self._label_worker.submit(job)

    def _join_labels(self):
        fut, self._label_future = self._label_future, None
        if fut is not None:
            fut.result()

    def close(self):
        """Wait for the background label draw and release the worker thread."""
        self._join_labels()
        if self._label_worker is not None:
            self._label_worker.shutdown(wait=True)
            self._label_worker = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_testing_seg(dataloader, dataset, model, criterion, logger, test_iter, writer, args):
    """The evaluation pass of the reference (``run_testing_seg``, utils/trainer.py:71-143) with the
    argmax, accuracy count and per-cloud part IoU on the device (``utils.metric.part_iou_from_logits``).

    The reference synchronises four times per batch (``.cpu()`` of the predictions, labels and
    one-hots, ``loss.item()``); here every batch only enqueues work and the B doubles / counters per
    batch are read back once after the loop.  Same return value: (accuracy, loss, mean per-category
    IoU, mean per-cloud IoU), the means taken in float64 on the host exactly as :119-122 does
    (``np.mean`` of an empty category is nan there as well).  ``logger`` / ``writer`` may be None."""
    import numpy as np
    from .utils.metric import object_names, part_iou_from_logits

    model.eval()
    ious, cats, corrects, losses = [], [], [], []
    for pts, cls, seg in dataloader:
        pts = pts.float().to(args.device)                                    # :91-94
        cls = cls.to(args.device)
        seg = seg.long().to(args.device)
        with torch.set_grad_enabled(False):                                  # :96-98
            pred, _ = model(pts, cls)
            losses.append(criterion(pred, seg).detach().double())
            iou, correct, cat = part_iou_from_logits(pred, seg, cls)
        ious.append(iou)
        cats.append(cat)
        corrects.append(correct.sum())
    if not ious:
        raise ValueError("run_testing_seg: empty dataloader")
    iou = torch.cat(ious).cpu().numpy()
    cat = torch.cat(cats).cpu().numpy()
    batch_correct = torch.stack(corrects).cpu().numpy()
    batch_loss = torch.stack(losses).cpu().numpy()

    total_accuracy = 0.0
    total_loss = 0.0
    for accu, loss in zip(batch_correct, batch_loss):                        # :101-103
        total_accuracy += accu / float(args.input_pts)
        total_loss += float(loss)
    shape_ious = [[] for _ in object_names]                                  # :86-88, :108-110
    for b in range(iou.shape[0]):
        shape_ious[int(cat[b])].append(iou[b])
    cat_per_iou = [np.mean(i) for i in shape_ious]                           # :119-122
    all_iou = [i for s in shape_ious for i in s]
    mean_cat_mious = np.mean(cat_per_iou)
    mean_all_mious = np.mean(all_iou)
    n = float(len(dataset))
    if logger is not None:
        logger.info("Test accuracy: {:.4f}\tloss: {:.3f}\tcat_iou: {:.4f}\tall_iou: {:.4f}".format(
            total_accuracy / n, total_loss / n, mean_cat_mious, mean_all_mious))
    if getattr(args, "tensorboard", False) and writer is not None:           # :133-141
        writer.add_scalar('Loss/test_cls', total_loss / n, test_iter)
        writer.add_scalar('Accuracy/test', total_accuracy / n, test_iter)
        writer.add_scalar('IoU/test_cat_iou', mean_cat_mious, test_iter)
        writer.add_scalar('IoU/test_all_iou', mean_all_mious, test_iter)
    return total_accuracy / n, total_loss / n, mean_cat_mious, mean_all_mious

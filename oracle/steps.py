"""CPU restatement of the train-step loop bodies in the reference's
``utils/trainer.py`` on top of the functional oracle models.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Also the ``port`` CPU
baseline timed by ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``).

A step takes parameter dicts of leaf tensors (requires_grad=True), runs the
same sequence of forwards / losses / backwards as the reference loop, and
leaves ``.grad`` on the leaves.  Optimizer updates are applied by the caller
(torch.optim.Adam, as train_segmentation.py:134-146 does).
"""
import torch
import torch.nn.functional as F

from . import pointnet_oracle as P
from . import discriminator_oracle as D


def make_D_label(shape, value, random, generator=None, device=None):
    """utils/utils.py:22-31: constant 0/1 target, or U(0, 0.305) for value 0 and
    U(0.7, 1.05) for value 1, drawn on the CPU (default generator there) and then
    moved to ``device`` (:31)."""
    if random:
        lo, hi = (0.0, 0.305) if value == 0 else (0.7, 1.05)
        lab = torch.empty(shape, dtype=torch.float32).uniform_(lo, hi, generator=generator)
    else:
        lab = torch.full(shape, float(value), dtype=torch.float32)
    return lab if device is None else lab.to(device)


def _set_requires_grad(params, flag):
    for p in params.values():
        p.requires_grad_(flag)


def adversarial_seg_step(g_params, d_params, batch_gt, batch_nogt, disc="pointwise",
                         lambda_seg=1.0, lambda_adv=1e-3, labels=None, generator=None,
                         branch=None, record=None):
    """One iteration of ``run_training_seg`` (utils/trainer.py:873-966) with
    history pools of size 0 (pass-through, utils/image_pool.py:35-36).

    G = PointNetSeg, D = PointwiseDiscNet (``disc="pointwise"``, the
    ``disc_seg`` factory mode, utils/model_utils.py:113-115) or ConvDiscNet
    (``disc="conv"``, the train_3D.py pairing -- it takes B x N x C, so the
    B x C x N maps are transposed first).

    ``labels``: optional (real_label, fake_label) tensors replacing the two
    random ``make_D_label`` draws at :940-945 and :955-960.
    ``branch`` / ``record``: optional dicts keyed by pass -- "g_gt", "g_nogt" (the two generator
    passes), "d_adv", "d_gt", "d_nogt" (the three discriminator passes) -- holding the per-pass
    ``branch`` / ``record`` dicts of the model oracles (branch-conditioned parity).
    Returns dict(l_seg, l_adv, l_D_gt, l_D_nogt) of Python floats.
    """
    pts, cls, seg = batch_gt
    pts_nogt, cls_nogt = batch_nogt
    n_pts = pts.shape[1]
    br = (lambda k: branch.get(k)) if branch is not None else (lambda k: None)
    rc = (lambda k: record.setdefault(k, {})) if record is not None else (lambda k: None)

    def run_D(x, which):
        if disc == "pointwise":
            return D.pointwise_disc_forward(d_params, x, n_pts, branch=br(which), record=rc(which))
        return D.conv_disc_forward(d_params, x.transpose(1, 2), branch=br(which), record=rc(which))

    # ---- train G (:884-929): D frozen
    _set_requires_grad(d_params, False)
    pred, _ = P.pointnet_seg_forward(g_params, pts, cls, branch=br("g_gt"), record=rc("g_gt"))   # :898
    l_seg = F.cross_entropy(pred, seg)                                        # :899
    pred_gt_softmax = F.softmax(pred, dim=1)                                  # :901
    pred_nogt, _ = P.pointnet_seg_forward(g_params, pts_nogt, cls_nogt, branch=br("g_nogt"),
                                          record=rc("g_nogt"))                # :913
    pred_nogt_softmax = F.log_softmax(pred_nogt, dim=1)                       # :914
    D_out = run_D(pred_nogt_softmax, "d_adv")                                 # :916
    l_adv = F.binary_cross_entropy_with_logits(D_out, make_D_label(D_out.shape, 1, False, device=D_out.device))
    (lambda_seg * l_seg + lambda_adv * l_adv).backward()                      # :927-929

    # ---- train D (:931-963)
    _set_requires_grad(d_params, True)
    D_out = run_D(pred_gt_softmax.detach(), "d_gt")                           # :936-938
    lab = labels[0] if labels is not None else make_D_label(D_out.shape, 1, True, generator, D_out.device)
    l_D_gt = F.binary_cross_entropy_with_logits(D_out, lab) * 0.5             # :946-947
    l_D_gt.backward()
    D_out = run_D(pred_nogt_softmax.detach(), "d_nogt")                       # :951-953
    lab = labels[1] if labels is not None else make_D_label(D_out.shape, 0, True, generator, D_out.device)
    l_D_nogt = F.binary_cross_entropy_with_logits(D_out, lab) * 0.5           # :961-962
    l_D_nogt.backward()
    return dict(l_seg=l_seg.item(), l_adv=l_adv.item(), l_D_gt=l_D_gt.item(),
                l_D_nogt=l_D_nogt.item())


def adversarial_cls_step(g_params, d_params, batch_gt, batch_nogt, lambda_cls=1.0,
                         lambda_adv=1e-3, labels=None, generator=None, training=False):
    """One iteration of ``run_training`` (utils/trainer.py:426-559): G =
    PointNetCls(40), D = DeepConvDiscNet(40, 1); both D inputs are log_softmax
    (:472, :492).  ``training=False`` evaluates PointNetCls without dropout so
    runs are repeatable."""
    pts, y = batch_gt
    (pts_nogt,) = batch_nogt
    _set_requires_grad(d_params, False)
    pred, _, _ = P.pointnet_cls_forward(g_params, pts, training=training)     # :468
    l_cls = F.cross_entropy(pred, y)
    pred_gt_ls = F.log_softmax(pred, dim=1)                                   # :472
    pred_nogt, _, _ = P.pointnet_cls_forward(g_params, pts_nogt, training=training)  # :490
    pred_nogt_ls = F.log_softmax(pred_nogt, dim=1)                            # :492
    D_out = D.deepconv_disc_forward(d_params, pred_nogt_ls)                   # :499
    l_adv = F.binary_cross_entropy_with_logits(D_out, make_D_label(D_out.shape, 1, False, device=D_out.device))
    (lambda_cls * l_cls + lambda_adv * l_adv).backward()
    _set_requires_grad(d_params, True)
    D_out = D.deepconv_disc_forward(d_params, pred_gt_ls.detach())            # :530
    lab = labels[0] if labels is not None else make_D_label(D_out.shape, 1, True, generator, D_out.device)
    l_D_gt = F.binary_cross_entropy_with_logits(D_out, lab) * 0.5
    l_D_gt.backward()
    D_out = D.deepconv_disc_forward(d_params, pred_nogt_ls.detach())          # :546
    lab = labels[1] if labels is not None else make_D_label(D_out.shape, 0, True, generator, D_out.device)
    l_D_nogt = F.binary_cross_entropy_with_logits(D_out, lab) * 0.5
    l_D_nogt.backward()
    return dict(l_cls=l_cls.item(), l_adv=l_adv.item(), l_D_gt=l_D_gt.item(),
                l_D_nogt=l_D_nogt.item())


def pointnet_cls_step(g_params, batch, feature_transform=False, lambda_cls=1.0,
                      lambda_regu=1e-3, training=False):
    """One iteration of ``run_training_pointnet_cls`` (utils/trainer.py:236-269):
    CE + lambda_regu * feature_transform_regularizer when trans_feat exists."""
    pts, y = batch
    pred, _, trans_feat = P.pointnet_cls_forward(g_params, pts, feature_transform,
                                                 training=training)
    l = F.cross_entropy(pred, y)
    out = dict(l_cls=l.item())
    if trans_feat is not None:
        l_regu = P.feature_transform_regularizer(trans_feat)
        out["l_regu"] = l_regu.item()
        loss = lambda_cls * l + lambda_regu * l_regu
    else:
        loss = lambda_cls * l
    loss.backward()
    return out


def _semi_labels(D_out, pred_nogt, semi_TH):
    """The pseudo-labels of the semi-supervised term (utils/trainer.py:727-739, :1998-2010): the
    generator's own argmax where the discriminator output exceeds ``semi_TH``, 255 (ignored) elsewhere.
    Returns (labels | None when every entry is ignored, fraction kept)."""
    mask = (D_out <= semi_TH).squeeze(1)
    semi_gt = torch.argmax(pred_nogt.detach(), dim=1)
    semi_gt[mask] = 255
    ratio = 1.0 - float(mask.sum().item()) / float(mask.numel())
    return (None if ratio == 0.0 else semi_gt), ratio


def adversarial_cls_semi_step(g_params, d_params, batch_gt, batch_nogt, optimizer, optimizer_D, i_iter,
                              semi_start, semi_TH, lambda_cls=1.0, lambda_adv=1e-3, lambda_semi=1.0,
                              labels=None, generator=None, training=False):
    """One iteration of ``run_training_semi`` (utils/trainer.py:635-794), optimizer calls included
    (zero_grad :644-645, steps :793-794): ``adversarial_cls_step`` plus, once
    ``i_iter > semi_start > 0``, CrossEntropyLoss(ignore_index=255) of the unlabelled logits against
    their own argmax where D_out > semi_TH (:727-739)."""
    optimizer.zero_grad(); optimizer_D.zero_grad()
    pts, y = batch_gt
    (pts_nogt,) = batch_nogt
    _set_requires_grad(d_params, False)
    pred, _, _ = P.pointnet_cls_forward(g_params, pts, training=training)     # :681
    l_cls = F.cross_entropy(pred, y)
    pred_gt_ls = F.log_softmax(pred, dim=1)                                   # :685
    pred_nogt, _, _ = P.pointnet_cls_forward(g_params, pts_nogt, training=training)  # :702
    pred_nogt_ls = F.log_softmax(pred_nogt, dim=1)                            # :704
    D_out = D.deepconv_disc_forward(d_params, pred_nogt_ls)                   # :711
    l_adv = F.binary_cross_entropy_with_logits(D_out, make_D_label(D_out.shape, 1, False))
    loss = lambda_cls * l_cls + lambda_adv * l_adv
    out = dict(l_cls=l_cls.item(), l_adv=l_adv.item(), l_semi=None)
    if semi_start > 0 and i_iter > semi_start:                                # :727
        semi_gt, _ = _semi_labels(D_out, pred_nogt, semi_TH)
        if semi_gt is not None:
            l_semi = F.cross_entropy(pred_nogt, semi_gt, ignore_index=255)    # semi_loss, train_classification.py:201
            loss = loss + lambda_semi * l_semi
            out["l_semi"] = l_semi.item()
    loss.backward()                                                           # :760
    _set_requires_grad(d_params, True)
    D_out = D.deepconv_disc_forward(d_params, pred_gt_ls.detach())            # :773
    lab = labels[0] if labels is not None else make_D_label(D_out.shape, 1, True, generator)
    l_D_gt = F.binary_cross_entropy_with_logits(D_out, lab) * 0.5
    l_D_gt.backward()
    D_out = D.deepconv_disc_forward(d_params, pred_nogt_ls.detach())          # :787
    lab = labels[1] if labels is not None else make_D_label(D_out.shape, 0, True, generator)
    l_D_nogt = F.binary_cross_entropy_with_logits(D_out, lab) * 0.5
    l_D_nogt.backward()
    optimizer.step(); optimizer_D.step()                                      # :793-794
    out.update(l_D_gt=l_D_gt.item(), l_D_nogt=l_D_nogt.item())
    return out


def adversarial_seg_semi_step(g_params, d_params, batch_gt, batch_nogt, optimizer, optimizer_D, i_iter,
                              semi_start, semi_TH, lambda_seg=1.0, lambda_adv=1e-3, lambda_semi=1.0,
                              labels=None, generator=None):
    """One iteration of ``run_training_seg_semi`` (utils/trainer.py:1927-2061), optimizer calls
    included: G = PointNetSeg, D = PointwiseDiscNet; the generator is stepped right after its backward
    (:2023), the discriminator at the end (:2061)."""
    optimizer.zero_grad(); optimizer_D.zero_grad()                            # :1937-1938
    pts, cls, seg = batch_gt
    pts_nogt, cls_nogt = batch_nogt
    n_pts = pts.shape[1]
    _set_requires_grad(d_params, False)
    pred, _ = P.pointnet_seg_forward(g_params, pts, cls)                      # :1969
    l_seg = F.cross_entropy(pred, seg)
    pred_gt_softmax = F.softmax(pred, dim=1)                                  # :1972
    pred_nogt, _ = P.pointnet_seg_forward(g_params, pts_nogt, cls_nogt)       # :1984
    pred_nogt_softmax = F.log_softmax(pred_nogt, dim=1)                       # :1985
    D_out = D.pointwise_disc_forward(d_params, pred_nogt_softmax, n_pts)      # :1987
    l_adv = F.binary_cross_entropy_with_logits(D_out, make_D_label(D_out.shape, 1, False))
    loss = lambda_seg * l_seg + lambda_adv * l_adv
    out = dict(l_seg=l_seg.item(), l_adv=l_adv.item(), l_semi=None)
    if semi_start > 0 and i_iter > semi_start:                                # :1999
        semi_gt, _ = _semi_labels(D_out, pred_nogt, semi_TH)
        if semi_gt is not None:
            l_semi = F.cross_entropy(pred_nogt, semi_gt, ignore_index=255)
            loss = loss + lambda_semi * l_semi
            out["l_semi"] = l_semi.item()
    loss.backward()
    optimizer.step()                                                          # :2022-2023
    _set_requires_grad(d_params, True)
    D_out = D.pointwise_disc_forward(d_params, pred_gt_softmax.detach(), n_pts)      # :2032
    lab = labels[0] if labels is not None else make_D_label(D_out.shape, 1, True, generator)
    l_D_gt = F.binary_cross_entropy_with_logits(D_out, lab) * 0.5
    l_D_gt.backward()
    D_out = D.pointwise_disc_forward(d_params, pred_nogt_softmax.detach(), n_pts)    # :2047
    lab = labels[1] if labels is not None else make_D_label(D_out.shape, 0, True, generator)
    l_D_nogt = F.binary_cross_entropy_with_logits(D_out, lab) * 0.5
    l_D_nogt.backward()
    optimizer_D.step()                                                        # :2061
    out.update(l_D_gt=l_D_gt.item(), l_D_nogt=l_D_nogt.item())
    return out


def adversarial_seg_dual_step(g_params, shared, shape, point, batch_gt, batch_nogt, optimizer,
                              optimizer_D_shape, optimizer_D_point, lambda_seg=1.0, lambda_adv=1e-3,
                              lambda_disc_shape=1.0, labels=None, generator=None):
    """One iteration of ``run_training_seg_dual`` (utils/trainer.py:2150-2284), optimizer calls
    included because they interleave with the passes: the generator steps right after its backward
    (:2225), ``optimizer_D_point`` (pointDisc + sharedDisc, train_segmentation.py:395-411) steps
    BEFORE the shape pass runs on the updated sharedDisc (:2273-2275), and -- as written --
    ``optimizer_D_shape.zero_grad()`` is called twice while ``optimizer_D_point`` is never zeroed
    (:2171-2172), so pointDisc's gradients accumulate over iterations and sharedDisc's point-phase
    gradient is still there when ``optimizer_D_shape`` steps.

    shared / shape / point: parameter dicts of BaseDiscNet / ShapeDiscNet / PointDiscNet."""
    optimizer.zero_grad()                                                     # :2170
    optimizer_D_shape.zero_grad()                                             # :2171
    optimizer_D_shape.zero_grad()                                             # :2172
    pts, cls, seg = batch_gt
    pts_nogt, cls_nogt = batch_nogt
    n_pts = pts.shape[1]
    for params in (shared, shape, point):                                     # :2174-2179
        _set_requires_grad(params, False)
    pred, _ = P.pointnet_seg_forward(g_params, pts, cls)                      # :2191
    l_seg = F.cross_entropy(pred, seg)
    pred_gt_softmax = F.softmax(pred, dim=1)                                  # :2194
    pred_nogt, _ = P.pointnet_seg_forward(g_params, pts_nogt, cls_nogt)       # :2206
    pred_nogt_softmax = F.log_softmax(pred_nogt, dim=1)                       # :2207
    D_point = D.point_disc_forward(point, D.base_disc_forward(shared, pred_nogt_softmax), n_pts)   # :2209-2210
    l_adv = F.binary_cross_entropy_with_logits(D_point, make_D_label(D_point.shape, 1, False))
    (lambda_seg * l_seg + lambda_adv * l_adv).backward()                      # :2222-2223
    optimizer.step()                                                          # :2224
    for params in (shared, shape, point):                                     # :2229-2234
        _set_requires_grad(params, True)
    D_point = D.point_disc_forward(point, D.base_disc_forward(shared, pred_gt_softmax.detach()), n_pts)
    lab = labels[0] if labels is not None else make_D_label(D_point.shape, 1, True, generator)
    l_D_point_gt = 0.5 * F.binary_cross_entropy_with_logits(D_point, lab)     # :2249
    D_point = D.point_disc_forward(point, D.base_disc_forward(shared, pred_nogt_softmax.detach()), n_pts)
    lab = labels[1] if labels is not None else make_D_label(D_point.shape, 0, True, generator)
    l_D_point_nogt = 0.5 * F.binary_cross_entropy_with_logits(D_point, lab)   # :2265
    (l_D_point_gt + l_D_point_nogt).backward()                                # :2273-2274
    optimizer_D_point.step()                                                  # :2275
    D_shape = D.shape_disc_forward(shape, D.base_disc_forward(shared, pred_gt_softmax.detach()))   # :2277-2278
    cls_gt = cls.argmax(dim=2).squeeze(1)                                     # :2279
    l_D_shape = F.cross_entropy(D_shape, cls_gt.long())                       # :2280
    (lambda_disc_shape * l_D_shape).backward()                                # :2283-2284
    optimizer_D_shape.step()                                                  # :2285
    return dict(l_seg=l_seg.item(), l_adv=l_adv.item(), l_D_point=l_D_point_gt.item() + l_D_point_nogt.item(),
                l_D_shape=l_D_shape.item())


def leaf_params(sd):
    """Detach-clone a state dict into autograd leaves."""
    return {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}

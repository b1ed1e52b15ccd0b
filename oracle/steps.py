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


def leaf_params(sd):
    """Detach-clone a state dict into autograd leaves."""
    return {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}

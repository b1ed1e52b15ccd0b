"""The adversarial train-step loop bodies of the reference's ``utils/trainer.py``
as callable steps on the libpcadv-backed modules.

The reference's ``run_training_seg`` (utils/trainer.py:849-1026) can drive the
modules unchanged; these functions are the same loop body (:873-966) without
the logging / checkpoint / evaluation code around it, so a benchmark or a
data-parallel launcher can call one iteration.  Losses come back as device
tensors: the caller decides when to synchronise (the reference calls ``.item()``
four times per iteration, :900, :925, :948, :963).
"""
import torch
import torch.nn.functional as F

from .utils.utils import make_D_label
from .utils.image_pool import ImagePool


def _device_label(like, value, random):
    """Device-side variant of make_D_label: same
    distributions, drawn by the device generator instead of the CPU one."""
    if not random:
        return torch.full_like(like, float(value))
    lo, hi = (0.0, 0.305) if value == 0 else (0.7, 1.05)
    return torch.empty_like(like).uniform_(lo, hi)


def adversarial_seg_step(model, model_D, gan_loss, seg_loss, optimizer, optimizer_D, batch_gt,
                         batch_nogt, args, history_pool_gt=None, history_pool_nogt=None,
                         device_labels=False):
    """One iteration of run_training_seg (utils/trainer.py:873-966).

    batch_gt = (pts B x N x 3, cls B x 1 x 16, seg B x N), batch_nogt = (pts, cls),
    already on ``args.device``.  ``args`` needs ``device``, ``lambda_seg``,
    ``lambda_adv``.  Returns (l_seg, l_adv, l_D) as 0-d device tensors."""
    gt_label, nogt_label = 1, 0
    pool_gt = history_pool_gt or ImagePool(0)
    pool_nogt = history_pool_nogt or ImagePool(0)

    def label(d_out, value, random):
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

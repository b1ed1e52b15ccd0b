"""Part-IoU evaluation with the reference's names (utils/metric.py), computed on the device.

The reference copies every prediction to the host and loops over clouds and parts in numpy
(utils/metric.py:20-39, called from utils/trainer.py:100-110).  Here the argmax, the per-part
intersection / union counts and the float64 IoU are two kernels of libpcadv
(``pcadv_part_counts`` / ``pcadv_part_iou``); only B doubles per batch ever need to leave the GPU.
"""
import torch

from .. import ops

# utils/metric.py:3-5 -- the ShapeNet-part categories in one-hot order and their part labels
object_names = ['Airplane', 'Bag', 'Cap', 'Car', 'Chair', 'Earphone', 'Guitar', 'Knife', 'Lamp',
                'Laptop', 'Motorbike', 'Mug', 'Pistol', 'Rocket', 'Skateboard', 'Table']
_PART_BEGIN = [0, 4, 6, 8, 12, 16, 19, 22, 24, 28, 30, 36, 38, 41, 44, 47, 50]
seg_classes = {name: list(range(_PART_BEGIN[i], _PART_BEGIN[i + 1])) for i, name in enumerate(object_names)}
seg_label_to_cat = {label: cat for cat, labels in seg_classes.items() for label in labels}

_BEGIN_CACHE = {}


def _part_begin(device):
    key = (device.type, device.index)
    if key not in _BEGIN_CACHE:
        _BEGIN_CACHE[key] = torch.tensor(_PART_BEGIN, dtype=torch.int32, device=device)
    return _BEGIN_CACHE[key]


def _onehot(batch_cls):
    cls = batch_cls[:, 0, :] if batch_cls.dim() == 3 else batch_cls          # utils/trainer.py:106
    if cls.dtype != torch.float32:
        cls = cls.float()
    return cls if cls.stride(-1) == 1 else cls.contiguous()


def part_iou_from_logits(pred, seg, cls, want_pred=False):
    """Device-side evaluation of one batch straight from the generator's output.

    pred: the B x C x N view ``PointNetSeg`` returns (point-major storage) or a B x N x C tensor;
    seg: int64 B x N; cls: B x 1 x 16 or B x 16 one-hot.
    Returns (iou float64 [B], correct int32 [B], category int32 [B][, pred_seg int64 B x N]) on the
    device, with no synchronisation: iou[b] is utils/metric.py:20-32's value for cloud b,
    correct[b] the cloud's share of utils/trainer.py:101."""
    B, N = seg.shape
    if pred.dim() != 3:
        raise ValueError("pred must be B x C x N or B x N x C")
    # B x N x C only when the shape says so unambiguously; otherwise the reference's B x C x N
    logits = pred if (pred.shape[1] == N and pred.shape[2] != N) else pred.transpose(1, 2)
    if tuple(logits.shape[:2]) != (B, N):
        raise ValueError("pred %s does not match seg %s" % (tuple(pred.shape), tuple(seg.shape)))
    if logits.stride(2) != 1:
        logits = logits.contiguous()
    counts, correct, out = ops.part_counts(seg, logits=logits.float(), want_pred=want_pred)
    iou, cat = ops.part_iou(counts, _onehot(cls), _part_begin(seg.device))
    return (iou, correct, cat, out) if want_pred else (iou, correct, cat)


def get_iou(gt, pred, cls_gt):
    """utils/metric.py:20-32 for one cloud of device tensors (gt, pred: int64 [N]; cls_gt: int)."""
    onehot = torch.zeros((1, len(object_names)), dtype=torch.float32, device=gt.device)
    onehot[0, int(cls_gt)] = 1.0
    counts, _, _ = ops.part_counts(gt[None], pred=pred[None])
    iou, _ = ops.part_iou(counts, onehot, _part_begin(gt.device))
    return float(iou.item())


def batch_get_iou(batch_pred, batch_seg, batch_cls):
    """utils/metric.py:34-39: per-cloud IoUs of int64 B x N predictions, as a list of floats
    (one device -> host copy of B doubles)."""
    counts, _, _ = ops.part_counts(batch_seg, pred=batch_pred)
    iou, _ = ops.part_iou(counts, _onehot(batch_cls), _part_begin(batch_seg.device))
    return iou.cpu().tolist()

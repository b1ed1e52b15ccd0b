"""Drop-in replacements for the reference's ``models/pointnet.py`` classes.

Same class names, constructor signatures, child-module names (``nn.Conv1d`` /
``nn.Linear`` parameter holders, so ``init_weights`` -- which selects modules by
class name, utils/model_utils.py:36-50 -- and old ``.pth`` state dicts keep
working), forward signatures, return tuples, shapes and strides.  ``forward``
does not run the child modules: it hands their parameters to an autograd
Function that calls the sm_100a kernels in libpcadv.so.  CUDA only.

Per-module arithmetic mode: ``model.precision = ops.Precision("fp32"|"fp16"|"bf16")``
(default: ``ops.default_precision()``).
"""
import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_NONE, ACT_RELU
from ._seg import SegFunction, PARAM_NAMES as _SEG_PARAMS
from ._mlp import point_mlp, BmmFunction, RegularizerFunction, LogSoftmaxRowsFunction

_RELU = (ACT_RELU, 0.0)
_NONE = (ACT_NONE, 0.0)


def _prec(module):
    return getattr(module, "precision", None) or ops.default_precision()


def _params(module, names):
    sd = dict(module.named_parameters())
    return [sd[n] for n in names]


class STN3d(nn.Module):
    """models/pointnet.py:14-43."""

    def __init__(self):
        super(STN3d, self).__init__()
        self.conv1 = torch.nn.Conv1d(3, 64, 1)
        self.conv2 = torch.nn.Conv1d(64, 128, 1)
        self.conv3 = torch.nn.Conv1d(128, 1024, 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, 9)
        self.relu = nn.ReLU()
        self.k = 3

    def _run(self, x_pm):
        """x_pm: point-major B x N x k (fp32).  Returns B x k x k."""
        B, N, k = x_pm.shape
        prec = _prec(self)
        g = point_mlp(prec, x_pm.reshape(B * N, k), [self.conv1, self.conv2, self.conv3],
                      [_RELU, _RELU, _RELU], reduce="points", group=N)            # B x 1024
        t = point_mlp(prec, g, [self.fc1, self.fc2, self.fc3], [_RELU, _RELU, _NONE])
        iden = torch.eye(k, dtype=t.dtype, device=t.device).reshape(1, k * k)
        return (t + iden).view(B, k, k)

    def forward(self, x):                       # B x k x N
        return self._run(x.transpose(1, 2))


class STNkd(STN3d):
    """models/pointnet.py:46-79."""

    def __init__(self, k=64):
        nn.Module.__init__(self)
        self.conv1 = torch.nn.Conv1d(k, 64, 1)
        self.conv2 = torch.nn.Conv1d(64, 128, 1)
        self.conv3 = torch.nn.Conv1d(128, 1024, 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k * k)
        self.relu = nn.ReLU()
        self.k = k


class PointNetfeat(nn.Module):
    """models/pointnet.py:81-137: 3->64->64 (ReLU), optional STNkd(64) + bmm,
    64->128 (ReLU), 128->1024 (no ReLU), max over points."""

    def __init__(self, global_feat=True, feature_transform=False):
        super(PointNetfeat, self).__init__()
        self.conv1 = torch.nn.Conv1d(3, 64, 1)
        self.conv2 = torch.nn.Conv1d(64, 64, 1)
        self.conv3 = torch.nn.Conv1d(64, 128, 1)
        self.conv4 = torch.nn.Conv1d(128, 1024, 1)
        self.global_feat = global_feat
        self.feature_transform = feature_transform
        if self.feature_transform:
            self.fstn = STNkd(k=64)

    def _run(self, x_pm, want_pointfeat=False):
        """x_pm: point-major B x N x 3.  Returns (global B x 1024, pointfeat
        B x N x 64 point-major fp32 | None, trans_feat | None)."""
        B, N, _ = x_pm.shape
        prec = _prec(self)
        x = x_pm.reshape(B * N, 3)
        if not self.feature_transform:
            convs = [self.conv1, self.conv2, self.conv3, self.conv4]
            acts = [_RELU, _RELU, _RELU, _NONE]
            if want_pointfeat:
                g, pf = point_mlp(prec, x, convs, acts, reduce="points", group=N, tap=1)
                return g, pf.view(B, N, 64), None
            return point_mlp(prec, x, convs, acts, reduce="points", group=N), None, None
        x2 = point_mlp(prec, x, [self.conv1, self.conv2], [_RELU, _RELU]).view(B, N, 64)
        trans_feat = self.fstn._run(x2)
        x2t = BmmFunction.apply(x2, trans_feat)
        g = point_mlp(prec, x2t.reshape(B * N, 64), [self.conv3, self.conv4], [_RELU, _NONE],
                      reduce="points", group=N)
        return g, x2t, trans_feat

    def forward(self, x):                       # B x 3 x N
        n_pts = x.size(2)
        g, pointfeat, trans_feat = self._run(x.transpose(1, 2), not self.global_feat)
        if self.global_feat:
            return g, trans_feat
        gt = g.view(-1, 1024, 1).repeat(1, 1, n_pts)
        return torch.cat([gt, pointfeat.transpose(1, 2)], 1), trans_feat


class PointNetCls(nn.Module):
    """models/pointnet.py:186-203."""

    def __init__(self, k=3, feature_transform=False):
        super(PointNetCls, self).__init__()
        self.feature_transform = feature_transform
        self.feat = PointNetfeat(global_feat=True, feature_transform=feature_transform)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k)
        self.dropout = nn.Dropout(p=0.3)
        self.relu = nn.ReLU()

    def forward(self, x):                       # B x N x 3
        g, _, trans_feat = self.feat._run(x)
        prec = _prec(self)
        if self.training:
            # fc2 -> Dropout(0.3) -> ReLU (:201).  The keep-mask is >= 0, so it commutes with
            # the ReLU: relu(drop(z)) = drop(relu(z)).
            h = point_mlp(prec, g, [self.fc1, self.fc2], [_RELU, _RELU])
            h = self.dropout(h)
            logits = point_mlp(prec, h, [self.fc3], [_NONE])
        else:
            logits = point_mlp(prec, g, [self.fc1, self.fc2, self.fc3], [_RELU, _RELU, _NONE])
        return logits, g.unsqueeze(2), trans_feat


class PointNetSeg(nn.Module):
    """models/pointnet.py:261-317."""

    def __init__(self, NUM_SEG_CLASSES):
        super(PointNetSeg, self).__init__()
        self.output_dim = NUM_SEG_CLASSES
        self.conv1 = torch.nn.Conv1d(3, 64, 1)
        self.conv2 = torch.nn.Conv1d(64, 128, 1)
        self.conv3 = torch.nn.Conv1d(128, 128, 1)
        self.conv4 = torch.nn.Conv1d(128, 128, 1)
        self.conv5 = torch.nn.Conv1d(128, 512, 1)
        self.conv6 = torch.nn.Conv1d(512, 2048, 1)
        self.fc1 = torch.nn.Linear(3024, 256)
        self.fc2 = torch.nn.Linear(256, 256)
        self.fc3 = torch.nn.Linear(256, 128)
        self.fc4 = torch.nn.Linear(128, self.output_dim)
        self._debug = None                      # tests set a dict here to capture saved tensors

    def forward(self, x, cls):                  # B x N x 3, B x 1 x 16
        logits, g = SegFunction.apply(_prec(self), self._debug, "logits", None, x, cls,
                                      *_params(self, _SEG_PARAMS))
        # B x k x N as a transposed view of point-major storage, as the reference
        # returns it (:315); B x 2048 x 1
        return logits.transpose(1, 2), g.unsqueeze(2)

    # ---- fused loss heads: the same forward with the trainer's softmax /
    # log_softmax / CrossEntropyLoss (utils/trainer.py:899-901, :914) folded into one pass over
    # the logits.  Optional: the reference-shaped forward() above stays the default.
    def forward_ce(self, x, cls, seg):
        """-> (CrossEntropyLoss(pred, seg) [0-d], softmax(pred) as a discriminator input,
        global B x 2048 x 1).  ``seg``: B x N int64 part labels."""
        loss, probs, g = SegFunction.apply(_prec(self), self._debug, "ce", seg, x, cls,
                                           *_params(self, _SEG_PARAMS))
        return loss, probs, g.unsqueeze(2)

    def forward_ce_logsoftmax(self, x, cls, seg, x_nogt, cls_nogt):
        """Both generator passes of one adversarial iteration (utils/trainer.py:898-901 and :913-914)
        as ONE pass over the labelled and the unlabelled clouds: -> (CrossEntropyLoss(pred, seg) of
        the labelled clouds, softmax(pred) of the labelled clouds [no grad], log_softmax(pred) of the
        unlabelled clouds [differentiable], global feature of all clouds (B + B') x 2048 x 1).  Clouds
        are independent through the whole network, so this equals forward_ce + forward_logsoftmax."""
        if x.shape[1] != x_nogt.shape[1]:
            raise ValueError("labelled and unlabelled clouds need the same number of points")
        loss, probs, lp, g = SegFunction.apply(_prec(self), self._debug, "ce+lsm", seg,
                                               torch.cat([x, x_nogt], 0), torch.cat([cls, cls_nogt], 0),
                                               *_params(self, _SEG_PARAMS))
        node = lp.grad_fn
        if node is not None and getattr(node, "box", None) is not None:
            lp._pcadv_box = node.box
        return loss, probs, lp, g.unsqueeze(2)

    def forward_logsoftmax(self, x, cls):
        """-> (log_softmax(pred, dim=1) as a differentiable discriminator input, global)."""
        lp, g = SegFunction.apply(_prec(self), self._debug, "lsm", None, x, cls,
                                  *_params(self, _SEG_PARAMS))
        node = lp.grad_fn
        if node is not None and getattr(node, "box", None) is not None:
            lp._pcadv_box = node.box
        return lp, g.unsqueeze(2)


class PointNetSeg_regulization(nn.Module):
    """models/pointnet.py:205-259: PointNetSeg with STN3d on the input and
    STNkd(128) after conv3."""

    def __init__(self, NUM_SEG_CLASSES):
        super(PointNetSeg_regulization, self).__init__()
        self.output_dim = NUM_SEG_CLASSES
        self.stn = STN3d()
        self.fstn = STNkd(k=128)
        self.conv1 = torch.nn.Conv1d(3, 64, 1)
        self.conv2 = torch.nn.Conv1d(64, 128, 1)
        self.conv3 = torch.nn.Conv1d(128, 128, 1)
        self.conv4 = torch.nn.Conv1d(128, 128, 1)
        self.conv5 = torch.nn.Conv1d(128, 512, 1)
        self.conv6 = torch.nn.Conv1d(512, 2048, 1)
        self.fc1 = torch.nn.Linear(3024, 256)
        self.fc2 = torch.nn.Linear(256, 256)
        self.fc3 = torch.nn.Linear(256, 128)
        self.fc4 = torch.nn.Linear(128, self.output_dim)

    def forward(self, x, cls):                  # B x N x 3, B x 1 x 16
        # Composition of the generic Functions (fp32 tensors between them, the T-Net transforms sit
        # between the trunk layers); the fused single-Function path is PointNetSeg's -- this variant
        # is not on the benchmarked path.
        B, N, _ = x.shape
        P = B * N
        prec = _prec(self)
        trans = self.stn._run(x)                                        # :229  B x 3 x 3
        xt = BmmFunction.apply(x, trans).reshape(P, 3)                  # :230-232
        x1 = point_mlp(prec, xt, [self.conv1], [_RELU])
        x2 = point_mlp(prec, x1, [self.conv2], [_RELU])
        x3 = point_mlp(prec, x2, [self.conv3], [_RELU])
        trans_feat = self.fstn._run(x3.view(B, N, 128))                 # :237  B x 128 x 128
        h = BmmFunction.apply(x3.view(B, N, 128), trans_feat).reshape(P, 128)   # :238-239
        x4 = point_mlp(prec, h, [self.conv4], [_RELU])
        x5 = point_mlp(prec, x4, [self.conv5], [_RELU])
        g = point_mlp(prec, x5, [self.conv6], [_RELU], reduce="points", group=N)   # B x 2048
        # concat fold (:246-251): global feature and class one-hot become a per-cloud bias
        w1 = self.fc1.weight
        cb = point_mlp(prec, torch.cat([g, cls.reshape(B, -1).float()], 1),
                       [(w1[:, 960:], self.fc1.bias)], [_NONE])          # B x 256
        # x1..x5 go in as five K-segments: the 960-channel map is not built either
        logits = point_mlp(prec, [x1, x2, x3, x4, x5],
                           [(w1[:, :960], None), self.fc2, self.fc3, self.fc4],
                           [_RELU, _RELU, _RELU, _NONE], group=N, group_bias=cb)   # P x k
        return logits.view(B, N, logits.shape[-1]).transpose(1, 2), g.unsqueeze(2), trans_feat


class PointNetDenseCls(nn.Module):
    """models/pointnet.py:320-343 with the two-line fix of SURVEY.md §8c-2 (the
    reference class raises as written): returns (log-probs B x N x k, trans_feat)."""

    def __init__(self, num_classes=16, feature_transform=False):
        super(PointNetDenseCls, self).__init__()
        self.num_classes = num_classes
        self.feature_transform = feature_transform
        self.feat = PointNetfeat(global_feat=False, feature_transform=feature_transform)
        self.conv1 = torch.nn.Conv1d(1088, 512, 1)
        self.conv2 = torch.nn.Conv1d(512, 256, 1)
        self.conv3 = torch.nn.Conv1d(256, 128, 1)
        self.conv4 = torch.nn.Conv1d(128, self.num_classes, 1)

    def forward(self, x):                       # B x 3 x N
        B, _, N = x.shape
        prec = _prec(self)
        g, pointfeat, trans_feat = self.feat._run(x.transpose(1, 2), True)
        # the 1088-channel concat [global(1024); pointfeat(64)] (:135-136) is never built:
        # conv1's global columns become a per-cloud bias, its pointfeat columns a K=64 layer
        w1 = self.conv1.weight.reshape(512, 1088)
        cb = point_mlp(prec, g, [(w1[:, :1024], self.conv1.bias)], [_NONE])       # B x 512
        h = point_mlp(prec, pointfeat.reshape(B * N, 64),
                      [(w1[:, 1024:], None), self.conv2, self.conv3, self.conv4],
                      [_RELU, _RELU, _RELU, _NONE], group=N, group_bias=cb)       # P x k
        h = LogSoftmaxRowsFunction.apply(h.view(-1, self.num_classes))             # :341
        return h.view(B, N, self.num_classes), trans_feat


def feature_transform_regularizer(trans):
    """models/pointnet.py:345-353: mean_b || T T^T - I ||_F (identity built on
    ``trans.device``)."""
    return RegularizerFunction.apply(trans)

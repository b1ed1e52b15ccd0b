"""Functional CPU restatement of the reference's ``models/pointnet.py``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each function takes a
state-dict ``sd`` with the reference's parameter names and plain tensors, and
uses ATen fp32 ops, so autograd supplies the reference backward.

Two hooks make *branch-conditioned* parity checks possible:

``branch``  optional dict.  ``branch["act:<layer>"]`` is a bool tensor with the
            shape of that layer's output (reference layout B x C x N, or B x C
            for per-cloud layers) that replaces the ReLU / LeakyReLU decision;
            ``branch["argmax:<layer>"]`` is an int64 B x C tensor that replaces
            the max-pool's argmax.  ReLU and max are the only non-smooth points
            of the network: with the decisions pinned the function is smooth
            and gradients of two implementations can be compared to rounding.
``record``  optional dict filled with the pre-activation tensors
            (``"pre:<layer>"``), the max-pool inputs (``"maxin:<layer>"``) and
            the oracle's own argmax (``"argmax:<layer>"``) so a test can show
            that every decision that differs sits within rounding distance of
            its boundary.
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# building blocks
# --------------------------------------------------------------------------- #
def _act(x, name, slope=0.0, branch=None, record=None):
    """ReLU (slope 0) or LeakyReLU; F.relu at models/pointnet.py:27-29,
    nn.LeakyReLU(0.2) at models/discriminator.py:39."""
    if record is not None:
        record["pre:" + name] = x
    if branch is not None and ("act:" + name) in branch:
        m = branch["act:" + name].to(x.dtype)
        return x * (m + (1.0 - m) * slope)
    if slope == 0.0:
        return F.relu(x)
    return F.leaky_relu(x, slope)


def _conv(sd, name, x):
    """nn.Conv1d(Cin, Cout, 1) on B x Cin x N (models/pointnet.py:17-19)."""
    return F.conv1d(x, sd[name + ".weight"], sd[name + ".bias"])


def _lin(sd, name, x):
    """nn.Linear (models/pointnet.py:20-22)."""
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _maxpool(x, name, branch=None, record=None):
    """torch.max(x, 2, keepdim=True)[0] over the point axis
    (models/pointnet.py:31, :64, :129, :303).  Returns B x C."""
    if record is not None:
        record["maxin:" + name] = x
        record["argmax:" + name] = x.max(2)[1]
    if branch is not None and ("argmax:" + name) in branch:
        idx = branch["argmax:" + name]
        return torch.gather(x, 2, idx.unsqueeze(2)).squeeze(2)
    return torch.max(x, 2, keepdim=True)[0].squeeze(2)


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------- #
# T-Nets
# --------------------------------------------------------------------------- #
def stn_forward(sd, x, k, tag="stn", branch=None, record=None):
    """STN3d (k=3, models/pointnet.py:14-43) and STNkd (:46-79): conv+ReLU
    k->64->128->1024, max over N, fc 1024->512->256->k*k, plus identity.
    x: B x k x N.  Returns B x k x k."""
    B = x.shape[0]
    h = _act(_conv(sd, "conv1", x), tag + ".conv1", 0.0, branch, record)
    h = _act(_conv(sd, "conv2", h), tag + ".conv2", 0.0, branch, record)
    h = _act(_conv(sd, "conv3", h), tag + ".conv3", 0.0, branch, record)
    g = _maxpool(h, tag + ".conv3", branch, record)            # B x 1024
    g = _act(_lin(sd, "fc1", g), tag + ".fc1", 0.0, branch, record)
    g = _act(_lin(sd, "fc2", g), tag + ".fc2", 0.0, branch, record)
    g = _lin(sd, "fc3", g)
    iden = torch.eye(k, dtype=x.dtype, device=x.device).reshape(1, k * k)
    return (g + iden).view(B, k, k)


def feature_transform_regularizer(trans):
    """models/pointnet.py:345-353 with the identity built on ``trans.device``
    instead of the hard-coded ``.cuda()`` at :349."""
    d = trans.size(1)
    eye = torch.eye(d, dtype=trans.dtype, device=trans.device)[None]
    return torch.mean(torch.norm(torch.bmm(trans, trans.transpose(2, 1)) - eye, dim=(1, 2)))


def feature_transform_mse(trans):
    """MSE orthogonality variant used by the stack trainers
    (utils/trainer.py:1276-1282): MSELoss(T T^T, I)."""
    d = trans.size(1)
    eye = torch.eye(d, dtype=trans.dtype, device=trans.device)[None].expand_as(
        torch.bmm(trans, trans.transpose(2, 1)))
    return F.mse_loss(torch.bmm(trans, trans.transpose(2, 1)), eye)


# --------------------------------------------------------------------------- #
# PointNetfeat / PointNetCls / PointNetDenseCls
# --------------------------------------------------------------------------- #
def pointnetfeat_forward(sd, x, global_feat=True, feature_transform=False,
                         branch=None, record=None):
    """PointNetfeat.forward, models/pointnet.py:109-136.  x: B x 3 x N.
    conv1/conv2 + ReLU, optional STNkd(64) + bmm, conv3 + ReLU, conv4 with NO
    ReLU (:128), max over N (:129)."""
    n_pts = x.shape[2]
    h = _act(_conv(sd, "conv1", x), "feat.conv1", 0.0, branch, record)
    h = _act(_conv(sd, "conv2", h), "feat.conv2", 0.0, branch, record)
    if feature_transform:
        trans_feat = stn_forward(_sub(sd, "fstn."), h, 64, "feat.fstn", branch, record)
        h = torch.bmm(h.transpose(2, 1), trans_feat).transpose(2, 1)
    else:
        trans_feat = None
    pointfeat = h
    h = _act(_conv(sd, "conv3", h), "feat.conv3", 0.0, branch, record)
    h = _conv(sd, "conv4", h)
    g = _maxpool(h, "feat.conv4", branch, record)               # B x 1024
    if global_feat:
        return g, trans_feat
    gt = g.view(-1, 1024, 1).repeat(1, 1, n_pts)
    return torch.cat([gt, pointfeat], 1), trans_feat


def pointnet_cls_forward(sd, pts, feature_transform=False, training=False,
                         dropout_mask=None, branch=None, record=None):
    """PointNetCls.forward, models/pointnet.py:197-203.  pts: B x N x 3.
    Returns (logits B x k, global B x 1024 x 1, trans_feat | None).
    Dropout(p=0.3) sits BEFORE the ReLU on fc2 (:201); in training mode pass
    ``dropout_mask`` (B x 256 keep-mask of 0/1) to make the run repeatable."""
    x = pts.transpose(1, 2)
    g, trans_feat = pointnetfeat_forward(_sub(sd, "feat."), x, True, feature_transform,
                                         branch, record)
    h = _act(_lin(sd, "fc1", g), "fc1", 0.0, branch, record)
    h = _lin(sd, "fc2", h)
    if training:
        if dropout_mask is None:
            h = F.dropout(h, 0.3, True)
        else:
            h = h * dropout_mask / (1.0 - 0.3)
    h = _act(h, "fc2", 0.0, branch, record)
    logits = _lin(sd, "fc3", h)
    return logits, g.unsqueeze(2), trans_feat


def pointnet_densecls_forward(sd, x, num_classes, feature_transform=False,
                              branch=None, record=None):
    """PointNetDenseCls.forward, models/pointnet.py:332-343, with the minimal
    fix of SURVEY.md §8c-2: ``x, trans_feat = self.feat(x)`` at :335 and
    ``self.num_classes`` for the undefined ``self.k`` at :341-342.  The class
    as written raises; this is the upstream fxia22/pointnet.pytorch behaviour
    the file header cites (:2).  x: B x 3 x N (no transpose, :333-335).
    Returns (log-probs B x N x k, trans_feat | None)."""
    B, _, n_pts = x.shape
    h, trans_feat = pointnetfeat_forward(_sub(sd, "feat."), x, False, feature_transform,
                                         branch, record)
    h = _act(_conv(sd, "conv1", h), "conv1", 0.0, branch, record)
    h = _act(_conv(sd, "conv2", h), "conv2", 0.0, branch, record)
    h = _act(_conv(sd, "conv3", h), "conv3", 0.0, branch, record)
    h = _conv(sd, "conv4", h)
    h = h.transpose(2, 1).contiguous()
    h = F.log_softmax(h.view(-1, num_classes), dim=-1)
    return h.view(B, n_pts, num_classes), trans_feat


# --------------------------------------------------------------------------- #
# PointNetSeg / PointNetSeg_regulization
# --------------------------------------------------------------------------- #
def pointnet_seg_forward(sd, pts, cls, regulization=False, branch=None, record=None):
    """PointNetSeg.forward (models/pointnet.py:282-317) and, with
    ``regulization=True``, PointNetSeg_regulization.forward (:226-259).

    pts: B x N x 3, cls: B x 1 x 16 one-hot.  conv+ReLU 3->64->128->128->128->
    512->2048 (:291-301), max over N (:303), tile global + class one-hot and
    concat to 3024 channels (:304-306), Linear+ReLU 3024->256->256->128 and
    Linear 128->k (:308-314).  Returns (logits B x k x N -- a transposed view of
    B x N x k storage, :315 -- , global B x 2048 x 1[, trans_feat B x 128 x 128])."""
    x = pts.transpose(1, 2)
    n_pts = x.shape[2]
    if regulization:
        trans = stn_forward(_sub(sd, "stn."), x, 3, "stn", branch, record)      # :229
        x = torch.bmm(x.transpose(2, 1), trans).transpose(2, 1)                 # :230-232
    x1 = _act(_conv(sd, "conv1", x), "conv1", 0.0, branch, record)
    x2 = _act(_conv(sd, "conv2", x1), "conv2", 0.0, branch, record)
    x3 = _act(_conv(sd, "conv3", x2), "conv3", 0.0, branch, record)
    if regulization:
        trans_feat = stn_forward(_sub(sd, "fstn."), x3, 128, "fstn", branch, record)  # :237
        h = torch.bmm(x3.transpose(2, 1), trans_feat).transpose(2, 1)                # :238-239
    else:
        h = x3
    x4 = _act(_conv(sd, "conv4", h), "conv4", 0.0, branch, record)
    x5 = _act(_conv(sd, "conv5", x4), "conv5", 0.0, branch, record)
    x6 = _act(_conv(sd, "conv6", x5), "conv6", 0.0, branch, record)
    g = _maxpool(x6, "conv6", branch, record)                                    # B x 2048
    x_global = g.unsqueeze(2)
    x_tile = x_global.repeat(1, 1, n_pts)
    cls_tile = cls.transpose(2, 1).repeat(1, 1, n_pts)
    x_all = torch.cat((x1, x2, x3, x4, x5, x_tile, cls_tile), 1)                # B x 3024 x N
    h = x_all.transpose(1, 2)
    h = _act(_lin(sd, "fc1", h), "fc1", 0.0, branch, record)
    h = _act(_lin(sd, "fc2", h), "fc2", 0.0, branch, record)
    h = _act(_lin(sd, "fc3", h), "fc3", 0.0, branch, record)
    h = _lin(sd, "fc4", h)
    out = h.transpose(1, 2)
    if regulization:
        return out, x_global, trans_feat
    return out, x_global


# --------------------------------------------------------------------------- #
# parameter construction (shapes = the reference's state-dict)
# --------------------------------------------------------------------------- #
def _conv_shapes(spec):
    out = {}
    for name, (cin, cout) in spec.items():
        out[name + ".weight"] = (cout, cin, 1)
        out[name + ".bias"] = (cout,)
    return out


def _lin_shapes(spec):
    out = {}
    for name, (cin, cout) in spec.items():
        out[name + ".weight"] = (cout, cin)
        out[name + ".bias"] = (cout,)
    return out


def stn_shapes(k):
    """models/pointnet.py:17-22 (k=3) and :49-54."""
    s = _conv_shapes({"conv1": (k, 64), "conv2": (64, 128), "conv3": (128, 1024)})
    s.update(_lin_shapes({"fc1": (1024, 512), "fc2": (512, 256), "fc3": (256, k * k)}))
    return s


def pointnet_seg_shapes(num_classes, regulization=False):
    """models/pointnet.py:268-278 (and :210-221 for the regulization variant)."""
    s = {}
    if regulization:
        s.update({"stn." + k: v for k, v in stn_shapes(3).items()})
        s.update({"fstn." + k: v for k, v in stn_shapes(128).items()})
    s.update(_conv_shapes({"conv1": (3, 64), "conv2": (64, 128), "conv3": (128, 128),
                           "conv4": (128, 128), "conv5": (128, 512), "conv6": (512, 2048)}))
    s.update(_lin_shapes({"fc1": (3024, 256), "fc2": (256, 256), "fc3": (256, 128),
                          "fc4": (128, num_classes)}))
    return s


def pointnet_cls_shapes(k, feature_transform=False):
    """models/pointnet.py:86-94 and :190-193."""
    s = {"feat." + n: v for n, v in _conv_shapes(
        {"conv1": (3, 64), "conv2": (64, 64), "conv3": (64, 128), "conv4": (128, 1024)}).items()}
    if feature_transform:
        s.update({"feat.fstn." + n: v for n, v in stn_shapes(64).items()})
    s.update(_lin_shapes({"fc1": (1024, 512), "fc2": (512, 256), "fc3": (256, k)}))
    return s


def pointnet_densecls_shapes(num_classes, feature_transform=False):
    """models/pointnet.py:325-329."""
    s = {"feat." + n: v for n, v in _conv_shapes(
        {"conv1": (3, 64), "conv2": (64, 64), "conv3": (64, 128), "conv4": (128, 1024)}).items()}
    if feature_transform:
        s.update({"feat.fstn." + n: v for n, v in stn_shapes(64).items()})
    s.update(_conv_shapes({"conv1": (1088, 512), "conv2": (512, 256), "conv3": (256, 128),
                           "conv4": (128, num_classes)}))
    return s


def xavier_state_dict(shapes, seed, gain=1.0):
    """Weights ~ xavier_normal_, biases 0: what ``init_weights(net, 'xavier')``
    produces (utils/model_utils.py:36-50).  Drawn from its own generator in
    state-dict order, so it does NOT reproduce ``torch.manual_seed`` + module
    construction; it is a convenience for tests that build both sides from one
    state dict."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        if name.endswith(".weight"):
            fan_out = shape[0] * (shape[2] if len(shape) == 3 else 1)
            fan_in = shape[1] * (shape[2] if len(shape) == 3 else 1)
            std = gain * (2.0 / (fan_in + fan_out)) ** 0.5
            sd[name] = torch.randn(shape, generator=g) * std
        else:
            sd[name] = torch.zeros(shape)
    return sd


def random_state_dict(shapes, seed, bias_scale=0.05):
    """Like ``xavier_state_dict`` but with non-zero biases, so bias paths are
    exercised by parity tests."""
    sd = xavier_state_dict(shapes, seed)
    g = torch.Generator().manual_seed(seed + 7919)
    for name in sd:
        if name.endswith(".bias"):
            sd[name] = torch.randn(sd[name].shape, generator=g) * bias_scale
    return sd


def synthetic_inputs(B, N, seed):
    """The synthetic batch of SURVEY.md §8c: drawn in this order from one
    generator -- pts U[-1,1)^3, y in [0,40), seg in [0,50), shape in [0,16)."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(B, N, 3, generator=g) * 2 - 1
    y = torch.randint(0, 40, (B,), generator=g)
    seg = torch.randint(0, 50, (B, N), generator=g)
    shp = torch.randint(0, 16, (B,), generator=g)
    cls = F.one_hot(shp, 16).to(torch.float32).view(B, 1, 16)
    return pts, y, seg, cls

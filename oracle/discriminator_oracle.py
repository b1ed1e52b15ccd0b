"""Functional CPU restatement of the reference's ``models/discriminator.py``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Same conventions as
``pointnet_oracle``: state-dict in, ATen fp32 ops, optional ``branch`` /
``record`` hooks.  ``rowargmax:<layer>`` pins the max over *channels* that
PointwiseDiscNet / PointDiscNet / StackDiscNet take per point.
"""
import torch
import torch.nn.functional as F

from .pointnet_oracle import _act, _conv, _lin, _maxpool, _conv_shapes, _lin_shapes


def _chanmax(x, name, branch=None, record=None):
    """max over the channel axis of B x C x N -> B x N
    (models/discriminator.py:70-72, :133-136, :169)."""
    if record is not None:
        record["chanmaxin:" + name] = x
        record["rowargmax:" + name] = x.max(1)[1]
    if branch is not None and ("rowargmax:" + name) in branch:
        idx = branch["rowargmax:" + name]
        return torch.gather(x, 1, idx.unsqueeze(1)).squeeze(1)
    return torch.max(x, 1)[0]


def conv_disc_forward(sd, x, branch=None, record=None):
    """ConvDiscNet.forward, models/discriminator.py:20-28.  x: B x N x C.
    conv+ReLU C->256->64->16, Linear 16->1 per point.  Returns B x N."""
    h = x.transpose(2, 1)
    h = _act(_conv(sd, "conv1", h), "conv1", 0.0, branch, record)
    h = _act(_conv(sd, "conv2", h), "conv2", 0.0, branch, record)
    h = _act(_conv(sd, "conv3", h), "conv3", 0.0, branch, record)
    h = _lin(sd, "fc", h.transpose(2, 1))
    return h.squeeze(2)


def deepconv_disc_forward(sd, x, branch=None, record=None):
    """DeepConvDiscNet.forward, models/discriminator.py:42-51.  x: B x C.
    conv+LeakyReLU(0.2) C->512->256->256->64->64 on B x C x 1, fc 64->out."""
    h = x.unsqueeze(2)
    for n in ("conv1", "conv2", "conv3", "conv4", "conv5"):
        h = _act(_conv(sd, n, h), n, 0.2, branch, record)
    return _lin(sd, "fc", h.view(-1, 64))


def pointwise_disc_forward(sd, x, input_pts, branch=None, record=None):
    """PointwiseDiscNet.forward, models/discriminator.py:62-79.  x: B x C x N.
    conv+ReLU C->64->64->64->128, max over channels.  Returns B x N."""
    h = x
    for n in ("conv1", "conv2", "conv3", "conv4"):
        h = _act(_conv(sd, n, h), n, 0.0, branch, record)
    return _chanmax(h, "conv4", branch, record).reshape(-1, input_pts)


def base_disc_forward(sd, x, branch=None, record=None, tag="base"):
    """BaseDiscNet.forward, models/discriminator.py:94-98: conv1..conv3 with
    LeakyReLU(0.2); conv4 exists in the state dict but is never applied."""
    h = x
    for n in ("conv1", "conv2", "conv3"):
        h = _act(_conv(sd, n, h), tag + "." + n, 0.2, branch, record)
    return h


def shape_disc_forward(sd, x, branch=None, record=None, tag="shape"):
    """ShapeDiscNet.forward, models/discriminator.py:111-117: conv 256->512 +
    LeakyReLU, max over N, fc 512->64 + LeakyReLU, fc 64->num_shapes."""
    h = _act(_conv(sd, "conv", x), tag + ".conv", 0.2, branch, record)
    g = _maxpool(h, tag + ".conv", branch, record)
    g = _act(_lin(sd, "fc1", g), tag + ".fc1", 0.2, branch, record)
    return _lin(sd, "fc2", g)


def point_disc_forward(sd, x, input_pts, branch=None, record=None, tag="point"):
    """PointDiscNet.forward, models/discriminator.py:130-137: conv+LeakyReLU
    256->256->128->128, max over channels.  Returns B x N."""
    h = x
    for n in ("conv1", "conv2", "conv3"):
        h = _act(_conv(sd, n, h), tag + "." + n, 0.2, branch, record)
    return _chanmax(h, tag + ".conv3", branch, record).reshape(-1, input_pts)


def stack_disc_forward(sd, x, branch=None, record=None):
    """StackDiscNet.forward, models/discriminator.py:162-175: conv+LeakyReLU(0.2)
    C->64->64->64->128, max over channels (keepdim) -> B x 1 x N, conv 1->S,
    custom activation z/(z+1) with z = logsumexp over S (:153-159).
    Returns (shape_logits B x S x N, disc_out B x N x 1)."""
    h = x
    for n in ("conv1", "conv2", "conv3", "conv4"):
        h = _act(_conv(sd, n, h), n, 0.2, branch, record)
    m = _chanmax(h, "conv4", branch, record).unsqueeze(1)
    shape_logits = _conv(sd, "conv5", m)
    z = torch.logsumexp(shape_logits.transpose(2, 1), dim=2, keepdim=True)
    return shape_logits, z / (z + 1.0)


# --------------------------------------------------------------------------- #
# state-dict shapes
# --------------------------------------------------------------------------- #
def conv_disc_shapes(input_dim):
    s = _conv_shapes({"conv1": (input_dim, 256), "conv2": (256, 64), "conv3": (64, 16)})
    s.update(_lin_shapes({"fc": (16, 1)}))
    return s


def deepconv_disc_shapes(input_dim, output_dim):
    s = _conv_shapes({"conv1": (input_dim, 512), "conv2": (512, 256), "conv3": (256, 256),
                      "conv4": (256, 64), "conv5": (64, 64)})
    s.update(_lin_shapes({"fc": (64, output_dim)}))
    return s


def pointwise_disc_shapes(input_dim):
    return _conv_shapes({"conv1": (input_dim, 64), "conv2": (64, 64), "conv3": (64, 64),
                         "conv4": (64, 128)})


def base_disc_shapes(input_dim, output_dim):
    return _conv_shapes({"conv1": (input_dim, 64), "conv2": (64, 64), "conv3": (64, output_dim),
                         "conv4": (output_dim, output_dim)})


def shape_disc_shapes(shared_output_dim, num_shapes):
    s = _conv_shapes({"conv": (shared_output_dim, 512)})
    s.update(_lin_shapes({"fc1": (512, 64), "fc2": (64, num_shapes)}))
    return s


def point_disc_shapes(shared_output_dim):
    return _conv_shapes({"conv1": (shared_output_dim, 256), "conv2": (256, 128),
                         "conv3": (128, 128)})


def stack_disc_shapes(input_dim, num_shapes):
    return _conv_shapes({"conv1": (input_dim, 64), "conv2": (64, 64), "conv3": (64, 64),
                         "conv4": (64, 128), "conv5": (1, num_shapes)})

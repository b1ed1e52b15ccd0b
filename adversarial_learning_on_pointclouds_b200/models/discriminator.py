"""Drop-in replacements for the reference's ``models/discriminator.py`` classes.

Same class names, constructor signatures, child-module names and return shapes;
``forward`` runs the layers through libpcadv (see ``models/pointnet.py``).
Inputs are the B x C x N maps the trainers feed them (softmax / log_softmax of
the segmentation logits, which are transposed views of point-major storage, so
``x.transpose(1, 2)`` is a free view) -- or B x N x C for ``ConvDiscNet`` and
B x C for ``DeepConvDiscNet``.
"""
import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_LEAKY, ACT_NONE, ACT_RELU
from ._mlp import point_mlp, LseRatioFunction

_RELU = (ACT_RELU, 0.0)
_NONE = (ACT_NONE, 0.0)
_LEAKY = (ACT_LEAKY, 0.2)


def _prec(module):
    return getattr(module, "precision", None) or ops.default_precision()


def _holders(module, *spec):
    """Register the reference's parameter holders in the reference's creation order (the order
    fixes which random numbers initialise which tensor): ``(name, cin, cout)`` makes the
    ``nn.Conv1d(cin, cout, 1)`` child ``name``; names starting with "fc" make ``nn.Linear``."""
    for name, cin, cout in spec:
        child = nn.Linear(cin, cout) if name.startswith("fc") else nn.Conv1d(cin, cout, 1)
        setattr(module, name, child)


def _leaky_holder(inplace=True):
    return nn.LeakyReLU(negative_slope=0.2, inplace=inplace)


def _rows(x_bcn):
    """The B x C x N map is handed to PointMLPFunction as is: it reads channel-major
    memory (torch softmax output) and transposed views of point-major storage (the
    generator's logits) without a torch-side copy.  A 16-bit B x N x Cpad tensor is the
    packed point-major map of the generator's fused softmax heads
    (PointNetSeg.forward_ce / forward_logsoftmax) and goes in as [B*N, Cpad] rows.
    Returns (rows, B, N, GradBox | None)."""
    if x_bcn.dim() == 3 and x_bcn.dtype in (torch.float16, torch.bfloat16):
        B, N, Cp = x_bcn.shape
        return x_bcn.reshape(B * N, Cp), B, N, getattr(x_bcn, "_pcadv_box", None)
    B, C, N = x_bcn.shape
    return x_bcn, B, N, None


class ConvDiscNet(nn.Module):
    """models/discriminator.py:10-28.  x: B x N x C -> B x N."""

    def __init__(self, input_dim):
        super().__init__()
        _holders(self, ("conv1", input_dim, 256), ("conv2", 256, 64), ("conv3", 64, 16), ("fc", 16, 1))
        self.relu = nn.ReLU()

    def forward(self, x):
        B, N, C = x.shape
        y = point_mlp(_prec(self), x.reshape(B * N, C), [self.conv1, self.conv2, self.conv3, self.fc],
                      [_RELU, _RELU, _RELU, _NONE])
        return y.view(B, N)


class DeepConvDiscNet(nn.Module):
    """models/discriminator.py:30-51.  x: B x C -> B x output_dim."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        widths = [input_dim, 512, 256, 256, 64, 64]
        _holders(self, *[("conv%d" % (i + 1), widths[i], widths[i + 1]) for i in range(5)],
                 ("fc", 64, output_dim))
        self.leaky_relu = _leaky_holder()

    def forward(self, x):
        return point_mlp(_prec(self), x, [self.conv1, self.conv2, self.conv3, self.conv4, self.conv5,
                                          self.fc], [_LEAKY] * 5 + [_NONE])


class PointwiseDiscNet(nn.Module):
    """models/discriminator.py:53-79.  x: B x C x N -> B x N (max over channels)."""

    def __init__(self, input_pts, input_dim):
        super().__init__()
        self.input_pts = input_pts
        widths = [input_dim, 64, 64, 64, 128]
        _holders(self, *[("conv%d" % (i + 1), widths[i], widths[i + 1]) for i in range(4)])

    def forward(self, x):
        rows, B, N, box = _rows(x)
        y = point_mlp(_prec(self), rows, [self.conv1, self.conv2, self.conv3, self.conv4],
                      [_RELU] * 4, reduce="channels", box=box)
        return y.view(-1, self.input_pts)


class BaseDiscNet(nn.Module):
    """models/discriminator.py:82-98 (conv4 is a parameter holder only: the
    reference's forward never applies it).  x: B x C x N -> B x output_dim x N."""

    def __init__(self, input_pts, input_dim, output_dim):
        super().__init__()
        self.input_pts = input_pts
        widths = [input_dim, 64, 64, output_dim, output_dim]
        _holders(self, *[("conv%d" % (i + 1), widths[i], widths[i + 1]) for i in range(4)])
        self.leaky_relu = _leaky_holder()

    def forward(self, x):
        rows, B, N, box = _rows(x)
        y = point_mlp(_prec(self), rows, [self.conv1, self.conv2, self.conv3], [_LEAKY] * 3, box=box)
        return y.view(B, N, y.shape[-1]).transpose(1, 2)


class ShapeDiscNet(nn.Module):
    """models/discriminator.py:100-117.  x: B x C x N -> B x num_shapes."""

    def __init__(self, shared_output_dim, num_shapes):
        super().__init__()
        self.interm_dim = 512
        _holders(self, ("conv", shared_output_dim, self.interm_dim), ("fc1", self.interm_dim, 64),
                 ("fc2", 64, num_shapes))
        self.leaky_relu = _leaky_holder()

    def forward(self, x):
        rows, B, N, _ = _rows(x)
        prec = _prec(self)
        g = point_mlp(prec, rows, [self.conv], [_LEAKY], reduce="points", group=N)   # B x 512
        return point_mlp(prec, g, [self.fc1, self.fc2], [_LEAKY, _NONE])


class PointDiscNet(nn.Module):
    """models/discriminator.py:119-137.  x: B x C x N -> B x N (max over channels)."""

    def __init__(self, shared_output_dim, input_pts):
        super().__init__()
        self.input_pts = input_pts
        _holders(self, ("conv1", shared_output_dim, 256), ("conv2", 256, 128), ("conv3", 128, 128))
        self.leaky_relu = _leaky_holder()

    def forward(self, x):
        rows, B, N, _ = _rows(x)
        y = point_mlp(_prec(self), rows, [self.conv1, self.conv2, self.conv3], [_LEAKY] * 3,
                      reduce="channels")
        return y.view(-1, self.input_pts)


class StackDiscNet(nn.Module):
    """models/discriminator.py:139-175.  x: B x C x N -> (shape_logits B x S x N,
    disc_out B x N x 1)."""

    def __init__(self, input_pts, input_dim, num_shapes):
        super().__init__()
        self.input_pts = input_pts
        widths = [input_dim, 64, 64, 64, 128]
        _holders(self, *[("conv%d" % (i + 1), widths[i], widths[i + 1]) for i in range(4)],
                 ("conv5", 1, num_shapes))
        self.leaky_relu = _leaky_holder(inplace=False)

    def custom_activation(self, x):
        """z / (z + 1) with z = logsumexp over the shape channel (models/discriminator.py:155-159);
        B x S x N -> B x N x 1."""
        z = torch.logsumexp(x.transpose(1, 2), dim=2, keepdim=True)
        return z / (z + 1.0)

    def forward(self, x):
        rows, B, N, box = _rows(x)
        prec = _prec(self)
        m = point_mlp(prec, rows, [self.conv1, self.conv2, self.conv3, self.conv4], [_LEAKY] * 4,
                      reduce="channels", box=box)                                 # [B*N]
        s = point_mlp(prec, m.view(B * N, 1), [self.conv5], [_NONE])              # [B*N, S]
        shape_logits = s.view(B, N, s.shape[-1]).transpose(1, 2)                           # B x S x N
        # custom_activation on the point-major rows (the same numbers as the B x S x N form above)
        return shape_logits, LseRatioFunction.apply(s).view(B, N, 1)

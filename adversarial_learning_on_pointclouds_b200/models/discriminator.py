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
from ._mlp import point_mlp

_RELU = (ACT_RELU, 0.0)
_NONE = (ACT_NONE, 0.0)
_LEAKY = (ACT_LEAKY, 0.2)


def _prec(module):
    return getattr(module, "precision", None) or ops.default_precision()


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
        super(ConvDiscNet, self).__init__()
        self.conv1 = torch.nn.Conv1d(input_dim, 256, 1)
        self.conv2 = torch.nn.Conv1d(256, 64, 1)
        self.conv3 = torch.nn.Conv1d(64, 16, 1)
        self.fc = nn.Linear(16, 1)
        self.relu = nn.ReLU()

    def forward(self, x):
        B, N, C = x.shape
        y = point_mlp(_prec(self), x.reshape(B * N, C), [self.conv1, self.conv2, self.conv3, self.fc],
                      [_RELU, _RELU, _RELU, _NONE])
        return y.view(B, N)


class DeepConvDiscNet(nn.Module):
    """models/discriminator.py:30-51.  x: B x C -> B x output_dim."""

    def __init__(self, input_dim, output_dim):
        super(DeepConvDiscNet, self).__init__()
        self.conv1 = torch.nn.Conv1d(input_dim, 512, 1)
        self.conv2 = torch.nn.Conv1d(512, 256, 1)
        self.conv3 = torch.nn.Conv1d(256, 256, 1)
        self.conv4 = torch.nn.Conv1d(256, 64, 1)
        self.conv5 = torch.nn.Conv1d(64, 64, 1)
        self.fc = nn.Linear(64, output_dim)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        return point_mlp(_prec(self), x, [self.conv1, self.conv2, self.conv3, self.conv4, self.conv5,
                                          self.fc], [_LEAKY] * 5 + [_NONE])


class PointwiseDiscNet(nn.Module):
    """models/discriminator.py:53-79.  x: B x C x N -> B x N (max over channels)."""

    def __init__(self, input_pts, input_dim):
        super(PointwiseDiscNet, self).__init__()
        self.input_pts = input_pts
        self.conv1 = torch.nn.Conv1d(input_dim, 64, 1)
        self.conv2 = torch.nn.Conv1d(64, 64, 1)
        self.conv3 = torch.nn.Conv1d(64, 64, 1)
        self.conv4 = torch.nn.Conv1d(64, 128, 1)

    def forward(self, x):
        rows, B, N, box = _rows(x)
        y = point_mlp(_prec(self), rows, [self.conv1, self.conv2, self.conv3, self.conv4],
                      [_RELU] * 4, reduce="channels", box=box)
        return y.view(-1, self.input_pts)


class BaseDiscNet(nn.Module):
    """models/discriminator.py:82-98 (conv4 is a parameter holder only: the
    reference's forward never applies it).  x: B x C x N -> B x output_dim x N."""

    def __init__(self, input_pts, input_dim, output_dim):
        super(BaseDiscNet, self).__init__()
        self.input_pts = input_pts
        self.conv1 = torch.nn.Conv1d(input_dim, 64, 1)
        self.conv2 = torch.nn.Conv1d(64, 64, 1)
        self.conv3 = torch.nn.Conv1d(64, output_dim, 1)
        self.conv4 = torch.nn.Conv1d(output_dim, output_dim, 1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        rows, B, N, box = _rows(x)
        y = point_mlp(_prec(self), rows, [self.conv1, self.conv2, self.conv3], [_LEAKY] * 3, box=box)
        return y.view(B, N, y.shape[-1]).transpose(1, 2)


class ShapeDiscNet(nn.Module):
    """models/discriminator.py:100-117.  x: B x C x N -> B x num_shapes."""

    def __init__(self, shared_output_dim, num_shapes):
        super(ShapeDiscNet, self).__init__()
        self.interm_dim = 512
        self.conv = torch.nn.Conv1d(shared_output_dim, self.interm_dim, 1)
        self.fc1 = torch.nn.Linear(self.interm_dim, 64)
        self.fc2 = torch.nn.Linear(64, num_shapes)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        rows, B, N, _ = _rows(x)
        prec = _prec(self)
        g = point_mlp(prec, rows, [self.conv], [_LEAKY], reduce="points", group=N)   # B x 512
        return point_mlp(prec, g, [self.fc1, self.fc2], [_LEAKY, _NONE])


class PointDiscNet(nn.Module):
    """models/discriminator.py:119-137.  x: B x C x N -> B x N (max over channels)."""

    def __init__(self, shared_output_dim, input_pts):
        super(PointDiscNet, self).__init__()
        self.input_pts = input_pts
        self.conv1 = torch.nn.Conv1d(shared_output_dim, 256, 1)
        self.conv2 = torch.nn.Conv1d(256, 128, 1)
        self.conv3 = torch.nn.Conv1d(128, 128, 1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        rows, B, N, _ = _rows(x)
        y = point_mlp(_prec(self), rows, [self.conv1, self.conv2, self.conv3], [_LEAKY] * 3,
                      reduce="channels")
        return y.view(-1, self.input_pts)


class StackDiscNet(nn.Module):
    """models/discriminator.py:139-175.  x: B x C x N -> (shape_logits B x S x N,
    disc_out B x N x 1)."""

    def __init__(self, input_pts, input_dim, num_shapes):
        super(StackDiscNet, self).__init__()
        self.input_pts = input_pts
        self.conv1 = torch.nn.Conv1d(input_dim, 64, 1)
        self.conv2 = torch.nn.Conv1d(64, 64, 1)
        self.conv3 = torch.nn.Conv1d(64, 64, 1)
        self.conv4 = torch.nn.Conv1d(64, 128, 1)
        self.conv5 = torch.nn.Conv1d(1, num_shapes, 1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2)

    def custom_activation(self, x):
        x = x.transpose(2, 1)
        out = torch.logsumexp(x, dim=2, keepdim=True)
        return out / (out + 1.0)

    def forward(self, x):
        rows, B, N, box = _rows(x)
        prec = _prec(self)
        m = point_mlp(prec, rows, [self.conv1, self.conv2, self.conv3, self.conv4], [_LEAKY] * 4,
                      reduce="channels", box=box)                                 # [B*N]
        s = point_mlp(prec, m.view(B * N, 1), [self.conv5], [_NONE])              # [B*N, S]
        shape_logits = s.view(B, N, s.shape[-1]).transpose(1, 2)                           # B x S x N
        return shape_logits, self.custom_activation(shape_logits)

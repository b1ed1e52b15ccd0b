"""OPTIONAL BatchNorm over point-major rows on libpcadv (``pcadv_bn_stats`` / ``_apply`` / ``_bwd``).

The reference has no BatchNorm anywhere (the only mentions are commented out,
models/pointnet.py:100-103, :160-163), so no drop-in module uses this layer and no parity run enables
it.  It exists because BASELINE.json's north_star words "BatchNorm statistics as warp-shuffle
reductions"; it follows ``torch.nn.BatchNorm1d`` (same parameters, buffers, momentum / eps semantics,
train / eval behaviour) and is tested against it.
"""
import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_NONE, ACT_RELU


class BatchNormRowsFunction(torch.autograd.Function):
    """(x [rows, C], gamma, beta) -> act(batchnorm(x)); statistics over the rows."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps, relu):
        if not x.is_cuda:
            raise RuntimeError("libpcadv layers need CUDA tensors; there is no CPU path")
        if x.dim() != 2 or not ops.bn_shape_ok(x.shape[1]):
            raise ValueError("BatchNormRows takes [rows, C] with C a power of two in [8, 2048]")
        if x.stride(1) != 1:
            x = x.contiguous()
        if training:
            if x.shape[0] < 2:
                raise ValueError("Expected more than 1 value per channel when training")
            mean, rstd = ops.bn_stats(x, eps, momentum, running_mean, running_var)
        else:
            mean, rstd = running_mean, torch.rsqrt(running_var + eps)
        y = ops.bn_apply(x, mean, rstd, gamma, beta, ACT_RELU if relu else ACT_NONE)
        ctx.relu, ctx.training = relu, training
        ctx.save_for_backward(x, mean, rstd, gamma, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma, y = ctx.saved_tensors
        if dy.stride(1) != 1:
            dy = dy.contiguous()
        if not ctx.training:
            # eval mode: the statistics are constants; y = x * (gamma * rstd) + const
            g = dy if y is None else dy * (y > 0)
            xhat = (x.float() - mean) * rstd
            dgamma, dbeta = (g.float() * xhat).sum(0), g.float().sum(0)
            dx = (g.float() * (rstd * (gamma if gamma is not None else 1.0))).to(dy.dtype)
            return dx, dgamma, dbeta, None, None, None, None, None, None
        dx, dgamma, dbeta = ops.bn_bwd(x, dy, mean, rstd, gamma, y=y, want_dx=ctx.needs_input_grad[0])
        return dx, dgamma, dbeta, None, None, None, None, None, None


class BatchNormRows(nn.Module):
    """``nn.BatchNorm1d(C)`` for point-major activations [rows, C] (rows = B * N points), with an
    optional fused ReLU.  Same state-dict keys as nn.BatchNorm1d."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, relu=False):
        super().__init__()
        self.num_features, self.eps, self.momentum, self.relu = num_features, eps, momentum, relu
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def forward(self, x):
        if self.training:
            self.num_batches_tracked += 1
        return BatchNormRowsFunction.apply(x, self.weight, self.bias, self.running_mean, self.running_var,
                                           self.training, self.momentum, self.eps, self.relu)


__all__ = ["BatchNormRows", "BatchNormRowsFunction"]

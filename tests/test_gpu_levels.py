"""GPU: the optional routings of the backward (pcadv_backlevel on the wide trunk levels / fc2 of
PointNetSeg, the one-hot level of the discriminators' max over channels) against the default routing.
Both sides are the same fp16 arithmetic with different summation orders, so the gradients agree far
inside the 1e-3 gate; the default routing itself is what tests/test_gpu_branch_parity.py checks
against the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from adversarial_learning_on_pointclouds_b200 import models as M
from adversarial_learning_on_pointclouds_b200.models import _mlp, _seg
from adversarial_learning_on_pointclouds_b200.ops import Precision
from adversarial_learning_on_pointclouds_b200.utils import init_net
from helpers import build_seg, inputs, randomize_biases, rel_err

DEV = "cuda"


def _seg_grads(levels, B, N):
    old = _seg._LEVELS
    _seg._LEVELS = frozenset(levels)
    try:
        net = build_seg(3, 11).to(DEV)
        net.precision = Precision("fp16")
        pts, _, seg, cls = inputs(B, N, 77)
        pred, g = net(pts.to(DEV), cls.to(DEV))
        loss = torch.nn.functional.cross_entropy(pred, seg.to(DEV)) + 1e-3 * g.square().mean()
        loss.backward()
        return loss.item(), {k: v.grad.clone() for k, v in net.named_parameters()}
    finally:
        _seg._LEVELS = old


@pytest.mark.parametrize("B,N", [(3, 384), (2, 1024)])
def test_seg_backward_levels_routings_agree(B, N):
    l0, g0 = _seg_grads([], B, N)
    for levels in (["fc4", "fc3"], ["fc4", "fc3", "fc2", "l4", "l3", "l2", "l1"], ["l3", "l1"]):
        l1, g1 = _seg_grads(levels, B, N)
        assert abs(l0 - l1) <= 1e-6 * abs(l0)
        worst = max(rel_err(g1[k], g0[k]) for k in g0)
        print(levels, "worst gradient difference against separate launches: %.2e" % worst)
        assert worst < 5e-4, (levels, worst)


@pytest.mark.parametrize("N", [700, 1024])
def test_disc_onehot_level_agrees_with_gather_kernels(N):
    def run(onehot):
        old = _mlp._ONEHOT_LEVEL
        _mlp._ONEHOT_LEVEL = onehot
        try:
            torch.manual_seed(5)
            d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
            randomize_biases([d], 13)
            d.to(DEV)
            d.precision = Precision("fp16")
            x = (torch.randn((3, 50, N), generator=torch.Generator().manual_seed(9)).softmax(1)).to(DEV)
            x.requires_grad_(True)
            out = d(x)
            (out * torch.linspace(0.5, 1.5, out.numel(), device=DEV).view_as(out)).mean().backward()
            return out.detach(), x.grad.clone(), {k: v.grad.clone() for k, v in d.named_parameters() if v.grad is not None}
        finally:
            _mlp._ONEHOT_LEVEL = old
    o0, dx0, g0 = run(False)
    o1, dx1, g1 = run(True)
    assert torch.equal(o0, o1)
    worst = max([rel_err(dx1, dx0)] + [rel_err(g1[k], g0[k]) for k in g0])
    print("PointwiseDiscNet", N, "worst difference one-hot level against gather kernels: %.2e" % worst)
    assert worst < 2e-3, worst


_PAIR_SNIPPET = r"""
import torch
from adversarial_learning_on_pointclouds_b200 import ops
from adversarial_learning_on_pointclouds_b200.ops import ACT_RELU, ENGINE_TC
torch.manual_seed(0)
worst = 0.0
for rows, ks, n in [(3000, [64, 128, 128, 128, 512], 256), (777, [512, 256], 128), (4133, [256], 512), (129, [128], 64)]:
    segs = [torch.randn(rows, k, device="cuda").half() for k in ks]
    w = (torch.randn(n, sum(ks), device="cuda") * 0.1).half()
    b = torch.randn(n, device="cuda")
    out, _, _ = ops.linear(segs, w, bias=b, act=ACT_RELU, out_dtype=torch.float16, engine=ENGINE_TC)
    ref = torch.relu(torch.cat(segs, 1).double() @ w.double().t() + b.double())
    worst = max(worst, ((out.double() - ref).norm() / ref.norm()).item())
print("PAIR_WORST %.3e" % worst)
"""


def test_rows_kernel_cta_pairs_with_multicast_weights():
    """The optional CTA-pair mode of the rows kernel (clusters of two, TMA-multicast weight tiles; the
    environment switch is read once per process, hence the child process), incl. an odd number of row
    tiles (the pair's phantom tail tile)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PCADV_ROWS_PAIR="2", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", _PAIR_SNIPPET], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    worst = float(r.stdout.strip().split("PAIR_WORST")[-1])
    print("rows kernel on CTA pairs: worst relative error %.2e" % worst)
    assert worst < 1e-3

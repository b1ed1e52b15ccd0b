"""Shared test helpers: seeded model builders that reproduce the recipes of
tests/golden/make_golden.py with the package's own modules, and comparison
utilities."""
import torch
import torch.nn.functional as F

from adversarial_learning_on_pointclouds_b200 import models as M
from adversarial_learning_on_pointclouds_b200.utils import init_net
from oracle.pointnet_oracle import synthetic_inputs


def probe_idx(numel, n=48, seed=99):
    g = torch.Generator().manual_seed(seed + numel)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def summarize(t):
    t = t.detach().to(torch.float32).cpu().contiguous().reshape(-1)
    return dict(norm=t.double().norm().item(), sum=t.double().sum().item(),
                abssum=t.double().abs().sum().item(), probe=t[probe_idx(t.numel())].clone(),
                numel=t.numel())


def check_weights(module, expected):
    """The module built here must carry exactly the weights the golden run had."""
    sd = module.state_dict()
    assert set(sd) == set(expected)
    for k, v in sd.items():
        s, a = expected[k]
        assert abs(v.double().sum().item() - s) <= 1e-9 * max(1.0, abs(s)), k
        assert abs(v.double().abs().sum().item() - a) <= 1e-9 * max(1.0, abs(a)), k


def randomize_biases(modules, seed):
    gb = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mm in modules:
            for n_, p in mm.named_parameters():
                if n_.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=gb) * 0.05)


def rel_err(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def assert_summary_close(t, ref, tol, what=""):
    """Compare a tensor with a stored summary (norm, sum of |.|, seeded probes)."""
    s = summarize(t)
    assert s["numel"] == ref["numel"], what
    scale = max(ref["norm"], 1e-30)
    assert abs(s["norm"] - ref["norm"]) <= tol * scale, (what, s["norm"], ref["norm"])
    pe = (s["probe"].double() - ref["probe"].double()).abs().max().item()
    pscale = max(ref["probe"].double().abs().max().item(), ref["norm"] / max(ref["numel"], 1) ** 0.5)
    assert pe <= 4 * tol * pscale, (what, pe, pscale)


def build_seg(wseed, bseed=None, regu=False):
    torch.manual_seed(wseed)
    net = init_net((M.PointNetSeg_regulization if regu else M.PointNetSeg)(50), "cpu", "xavier")
    if bseed is not None:
        randomize_biases([net], bseed)
    return net


inputs = synthetic_inputs
__all__ = ["F"]

"""Auxiliary kernels: on-device jitter, the optional BatchNorm layer
(north-star wording), workspace queries, and the input pipeline of the graphed step.
CPU part: the jitter oracle's stream (known answers, moments); GPU part: kernels against the oracle /
torch.nn.BatchNorm1d."""
import numpy as np
import pytest
import torch

from oracle import jitter_oracle as JO


def test_philox_known_answers():
    """Philox-4x32-10 known-answer vectors (Random123 kat_vectors): counter 0 / key 0, and the
    all-ones pattern."""
    z = JO.philox4x32_10(np.array([0], dtype=np.uint64), 0)[0]
    assert [int(v) for v in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


def test_jitter_oracle_matches_the_reference_distribution():
    """The oracle stream has the moments of clip(sigma * N(0, 1)) that the reference's numpy jitter
    has (dataset/modelNetData.py:88): zero mean, the clipped standard deviation, hard bounds."""
    data = np.zeros((200000, 3), dtype=np.float32)
    mine = JO.jitter(data, sigma=0.02, clip=0.05, seed=123)
    ref = JO.reference_jitter(data.astype(np.float64), sigma=0.02, clip=0.05, rng=np.random.RandomState(5))
    assert np.abs(mine).max() <= 0.05 + 1e-7 and np.abs(ref).max() <= 0.05 + 1e-12
    assert abs(mine.mean()) < 2e-4 and abs(mine.std() - ref.std()) < 2e-4
    # fraction clipped: P(|z| > 2.5) = 1.24 %
    assert abs((np.abs(mine) >= 0.05 - 1e-7).mean() - 0.0124) < 2e-3
    # independent of how the stream is cut: an offset continues it
    a = JO.normals(4096, 7)
    b = np.concatenate([JO.normals(1024, 7), JO.normals(3072, 7, offset=256)])
    assert np.array_equal(a, b)


gpu = pytest.mark.gpu


@gpu
def test_jitter_kernel_matches_oracle_stream():
    from adversarial_learning_on_pointclouds_b200 import ops
    g = torch.Generator().manual_seed(3)
    pts = (torch.rand(37, 1001, 3, generator=g) * 2 - 1)
    out = ops.jitter(pts.cuda(), sigma=0.01, clip=0.05, seed=0x1234567890ABCDEF, offset=11)
    want = JO.jitter(pts.numpy(), 0.01, 0.05, seed=0x1234567890ABCDEF, offset=11)
    # same counter-based stream; the device's logf / sincosf differ from numpy's in the last ulps
    assert np.abs(out.cpu().numpy() - want).max() < 2e-7
    d = (out.cpu() - pts)
    assert d.abs().max().item() <= 0.05 + 1e-7
    assert abs(d.mean().item()) < 1e-4 and abs(d.std().item() - 0.01) < 1e-4
    # in place, and a different seed gives a different draw
    p2 = pts.cuda().clone()
    ops.jitter(p2, seed=0x1234567890ABCDEF, offset=11, out=p2)
    assert torch.equal(p2, out)
    assert not torch.equal(ops.jitter(pts.cuda(), seed=5), ops.jitter(pts.cuda(), seed=6))
    assert ops.jitter(torch.zeros(0, 3, device="cuda")).numel() == 0
    with pytest.raises(ValueError):
        ops.jitter(pts.cuda(), clip=0.0)


@gpu
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("rows,C,dtype", [(1000, 64, torch.float32), (4097, 128, torch.float16), (333, 1024, torch.float32),
                                           (5000, 8, torch.float32), (16, 256, torch.float32)])
def test_batchnorm_rows_matches_torch(rows, C, dtype, relu):
    """The optional BatchNorm layer against torch.nn.BatchNorm1d (+ ReLU): output, running statistics,
    input / weight / bias gradients, train and eval mode.  A large common offset checks the pivoted
    variance."""
    from adversarial_learning_on_pointclouds_b200.models._bn import BatchNormRows
    g = torch.Generator().manual_seed(rows + C)
    x0 = (torch.randn(rows, C, generator=g) * 0.7 + 30.0 * torch.randn(1, C, generator=g)).to(dtype)
    w = torch.randn(rows, C, generator=g)
    mine = BatchNormRows(C, relu=relu).cuda()
    ref = torch.nn.BatchNorm1d(C).cuda().double()
    with torch.no_grad():
        mine.weight.copy_(torch.rand(C, generator=g) + 0.5); mine.bias.copy_(torch.randn(C, generator=g) * 0.1)
        ref.weight.copy_(mine.weight.double()); ref.bias.copy_(mine.bias.double())
    tol = 2e-5 if dtype == torch.float32 else 2e-3
    for training in (True, False):
        mine.train(training); ref.train(training)
        x = x0.clone().cuda().requires_grad_(True)
        xr = x0.double().cuda().requires_grad_(True)
        y = mine(x)
        yr = ref(xr)
        if relu:
            yr = torch.relu(yr)
        (y.float() * w.cuda()).sum().backward()
        (yr * w.double().cuda()).sum().backward()
        err = lambda a, b: ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()
        assert err(y, yr) < tol, (training, err(y, yr))
        assert err(x.grad, xr.grad) < 5 * tol, (training, err(x.grad, xr.grad))
        assert err(mine.weight.grad, ref.weight.grad) < 5 * tol and err(mine.bias.grad, ref.bias.grad) < 5 * tol
        assert err(mine.running_mean, ref.running_mean) < tol and err(mine.running_var, ref.running_var) < 10 * tol
        mine.zero_grad(); ref.zero_grad()
    assert int(mine.num_batches_tracked) == 1


def test_query_workspace_matches_the_documented_sizes():
    from adversarial_learning_on_pointclouds_b200 import ops, _lib
    assert ops.query_workspace(_lib.WS_MAXPOOL_BWD_INPLACE, groups=256, rows_per_group=4096, n=2048) == \
        256 * (4096 + 3 * 2048 + 4 * 2048) * 4
    assert ops.query_workspace(_lib.WS_MAXPOOL_BWD_INPLACE, groups=3, rows_per_group=201, n=1024) % 16 == 0
    assert ops.query_workspace(_lib.WS_AMAX_SCALE) == 4
    with pytest.raises(_lib.PcadvError):
        ops.query_workspace(99)

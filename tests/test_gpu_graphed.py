"""Parity of the object bench.py times -- ``GraphedAdversarialSegStep(fused=True)``: CUDA-graph
replay from static buffers, pinned-label upload, prefetching input pipeline, capturable fused
Adam -- against the eager loop body run from the same state and against ``oracle.steps``
(utils/trainer.py:873-966).  Plus the step-level regressions of the round-1 review: history pools
of size > 0, the forward cache inside a step scope, ignored labels in the fused CE head, and the
one-pass generator against the two-pass one."""
import argparse
import copy
import os
import random
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from adversarial_learning_on_pointclouds_b200 import models as M, Precision            # noqa: E402
from adversarial_learning_on_pointclouds_b200.models._chain import weight_cache          # noqa: E402
from adversarial_learning_on_pointclouds_b200.trainer import (adversarial_seg_step,     # noqa: E402
                                                              adversarial_seg_step_fused,
                                                              GraphedAdversarialSegStep)
from adversarial_learning_on_pointclouds_b200.utils import init_net, ImagePool          # noqa: E402
from oracle import steps                                                                 # noqa: E402
from helpers import inputs, randomize_biases, rel_err, build_seg                         # noqa: E402

DEV = torch.device("cuda", 0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _models(N, mode, seed=1):
    torch.manual_seed(seed)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
    randomize_biases([g, d], 3)
    g.precision = d.precision = Precision(mode)
    return g, d


def _adam(g, d, capturable, optim="adam"):
    if optim == "sgd":
        # plain SGD: a decision that lands on the other side moves the parameters by lr x (1e-3 of the
        # gradient), not by an lr-sized step per affected entry as under Adam -- the sensitive variant
        return (torch.optim.SGD(g.parameters(), lr=0.2, fused=True), torch.optim.SGD(d.parameters(), lr=0.02, fused=True))
    return (torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999), fused=True, capturable=capturable),
            torch.optim.Adam(d.parameters(), lr=1e-5, betas=(0.9, 0.999), fused=True, capturable=capturable))


def _batches(B, N, n_iter):
    out = []
    for it in range(n_iter):
        pts, _, seg, cls = inputs(B, N, 100 + it)
        pts2, _, _, cls2 = inputs(B, N, 500 + it)
        out.append(((pts, cls, seg), (pts2, cls2)))
    return out


@pytest.mark.parametrize("optim", ["adam", "sgd"])
@pytest.mark.parametrize("one_pass", [True, False])
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_graphed_fused_step_matches_eager_and_oracle(mode, one_pass, optim):
    """>= 5 replays with a fresh batch each (prefetch / step_prefetched from pinned host memory)
    against (a) the eager ``adversarial_seg_step_fused`` from the same initial state: losses and
    every G / D parameter; (b) in the fp32 mode, ``oracle.steps`` + Adam on the CPU."""
    B, N, iters = 3, 384, 6
    g, d = _models(N, mode)
    g2, d2 = copy.deepcopy(g), copy.deepcopy(d)
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    start = {k: v.clone() for k, v in g.state_dict().items()}
    g.to(DEV); d.to(DEV); g2.to(DEV); d2.to(DEV)
    opt, optD = _adam(g, d, True, optim)
    opt2, optD2 = _adam(g2, d2, True, optim)
    if optim == "sgd":
        ropt, roptD = torch.optim.SGD(list(gp.values()), lr=0.2), torch.optim.SGD(list(dp.values()), lr=0.02)
    else:
        ropt = torch.optim.Adam(list(gp.values()), lr=1e-4, betas=(0.9, 0.999))
        roptD = torch.optim.Adam(list(dp.values()), lr=1e-5, betas=(0.9, 0.999))
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=1e-3)
    gan, ce = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()
    batches = _batches(B, N, iters)
    pinned = [tuple(tuple(t.pin_memory() for t in part) for part in b) for b in batches]
    on_dev = [tuple(tuple(t.to(DEV) for t in part) for part in b) for b in batches]

    # ---- graphed arm: label draws come from the CPU generator, iteration by iteration
    torch.manual_seed(4242)
    gstep = GraphedAdversarialSegStep(g, d, gan, ce, opt, optD, targs, on_dev[0][0], on_dev[0][1],
                                      warmup=2, fused=True, one_pass=one_pass)
    # construction (warm-up + capture) must not have trained anything
    for k, v in g.state_dict().items():
        assert torch.equal(v.cpu(), start[k]), k
    got = []
    gstep.prefetch(*pinned[0])
    for it in range(iters):
        losses = gstep.step_prefetched()
        if it + 1 < iters:
            gstep.prefetch(*pinned[it + 1])
        got.append(losses.clone())
    torch.cuda.synchronize()
    gstep.close()
    got = [t.cpu() for t in got]

    # ---- eager arm, same state, same label stream
    torch.manual_seed(4242)
    want = []
    for it in range(iters):
        l = adversarial_seg_step_fused(g2, d2, gan, ce, opt2, optD2, on_dev[it][0], on_dev[it][1], targs,
                                       one_pass=one_pass)
        want.append(torch.stack(l).cpu())
    # Same kernels, same inputs: the two arms differ only by the order of the fp32 atomics in the
    # weight-gradient kernels (~1e-7), which now and then lands one ReLU / argmax decision of a later
    # iteration on the other side (the adversarial loss then moves by a few 1e-5); in the fp16 mode the
    # 16-bit engine copy of a weight is a step function of the fp32 master, so the noise also moves
    # copies by one fp16 ulp from the second iteration on.
    ltol = 2e-4 if mode == "fp32" else 1e-3
    worst_loss = max(((got[it] - want[it]).abs() / want[it].abs()).max().item() for it in range(iters))
    for it in range(iters):
        assert torch.allclose(got[it], want[it], rtol=ltol, atol=1e-7), (it, got[it].tolist(), want[it].tolist())
    worst = 0.0
    pairs = list(zip(list(g.named_parameters()) + list(d.named_parameters()),
                     list(g2.named_parameters()) + list(d2.named_parameters())))
    for (k, a), (_, b) in pairs:
        e = rel_err(a, b)
        worst = max(worst, e)
        if mode == "fp32":
            assert e < 1e-3, (k, e, optim)
    moved_g = torch.cat([(v.detach().cpu() - start[k]).flatten() for k, v in g.named_parameters()]).norm().item()
    apart_g = torch.cat([(a.detach() - b.detach()).flatten() for (k, a), (_, b) in pairs[:20]]).norm().item()
    print("graph vs eager (%s): worst loss rel diff over %d iterations %.2e; G parameters moved %.3e, apart %.3e"
          % (mode, iters, worst_loss, moved_g, apart_g))
    # Adam: 2e-6 (fp32) / 1-3e-2 (fp16) of the distance moved, but a ReLU / argmax decision of some
    # iteration that lands on the other side (atomics order, ~1e-7) is turned into lr-sized steps by
    # Adam's sign-like updates: seen up to 2.5e-3 (fp32) and > 5e-2 (fp16), so those bounds are loose.
    # SGD does not amplify it: the bounds there are what catches a real fault -- the cross-stream race
    # this test caught while the discriminator phase moved to a second stream showed as 8.5e-2.
    if optim == "sgd":
        bound = 2e-3 if mode == "fp32" else 8e-2      # measured 3e-6 / 1.7-3.5e-2
    else:
        bound = 5e-2 if mode == "fp32" else 0.15      # measured 2e-6 (2.2e-2 with a flipped decision) / 1-3e-2
    assert apart_g < bound * moved_g, (apart_g, moved_g, optim)
    print("graph vs eager (%s, one_pass=%s, %s): worst parameter rel err after %d steps %.2e"
          % (mode, one_pass, optim, iters, worst))
    assert gstep.launches_per_step > 0

    # ---- oracle arm (fp32 verification mode): same batches, same label stream
    if mode == "fp32":
        torch.manual_seed(4242)
        for it in range(iters):
            ropt.zero_grad(); roptD.zero_grad()
            r = steps.adversarial_seg_step(gp, dp, batches[it][0], batches[it][1])
            ropt.step(); roptD.step()
            o = torch.tensor([r["l_seg"], r["l_adv"], r["l_D_gt"] + r["l_D_nogt"]])
            assert torch.allclose(got[it], o, rtol=2e-4, atol=1e-6), (it, got[it], o)
        moved = torch.cat([(gp[k].detach() - start[k]).flatten() for k in gp]).norm().item()
        apart = torch.cat([(v.detach().cpu() - gp[k].detach()).flatten() for k, v in g.named_parameters()]).norm().item()
        print("graph vs oracle: parameters moved %.3e, apart %.3e" % (moved, apart))
        # Adam turns rounding-level differences of near-zero gradient entries into lr-sized steps
        # (DESIGN.md 5); the bound is relative to how far training moved the parameters
        assert apart < 0.05 * moved, (apart, moved)


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_one_pass_generator_equals_two_pass(mode):
    """forward_ce_logsoftmax (one pass over labelled + unlabelled clouds, one common gradient
    scale) against forward_ce + forward_logsoftmax: losses, discriminator inputs, all gradients."""
    B, N = 3, 512              # > 1024 rows per pass: both variants run the same kernels per row
    g, d = _models(N, mode, seed=7)
    g.to(DEV); d.to(DEV)
    (pts, cls, seg), (pts2, cls2) = [tuple(t.to(DEV) for t in part) for part in _batches(B, N, 1)[0]]
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=0.5)
    gan, ce = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()
    res = {}
    for one_pass in (False, True):
        opt = torch.optim.SGD(g.parameters(), lr=0.0)
        optD = torch.optim.SGD(d.parameters(), lr=0.0)
        torch.manual_seed(5)
        l = adversarial_seg_step_fused(g, d, gan, ce, opt, optD, (pts, cls, seg), (pts2, cls2), targs,
                                       one_pass=one_pass)
        res[one_pass] = (torch.stack(l).cpu(), {k: v.grad.clone() for k, v in
                                                list(g.named_parameters()) + list(d.named_parameters())})
    tol = 1e-5 if mode == "fp32" else 1e-3
    assert torch.allclose(res[True][0], res[False][0], rtol=tol, atol=1e-6)
    errs = {k: rel_err(res[True][1][k], res[False][1][k]) for k in res[True][1]}
    print("one-pass vs two-pass (%s): worst gradient rel err %.2e" % (mode, max(errs.values())))
    # fp32: identical per-row arithmetic, the sums over rows differ in order only.  fp16: the forward is
    # bit-identical, but the adversarial rows of the backward carry a different gradient scale, so every
    # fp16 rounding of dz differs: the two runs are two independent draws of the mode's rounding noise
    # around the oracle (tests/test_gpu_branch_parity.py); measured 6e-5 here.
    assert max(errs.values()) < (2e-5 if mode == "fp32" else 2e-3), errs


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_history_pools_with_positive_size(mode, fused):
    """pool_size > 0 (utils/image_pool.py:26-55): the pool hands the discriminator a fresh leaf
    (a clone, possibly of an older sample); its gradient is discarded by the reference.  Fused and
    reference-shaped loop bodies must both run and agree on the discriminator's gradients."""
    B, N = 3, 256
    g, d = _models(N, mode, seed=9)
    g.to(DEV); d.to(DEV)
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=0.5)
    gan, ce = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()
    grads = {}
    for arm in (fused, None):                                    # None: pool-free control run
        opt = torch.optim.SGD(g.parameters(), lr=0.0)
        optD = torch.optim.SGD(d.parameters(), lr=0.0)
        pools = (ImagePool(4), ImagePool(4)) if arm is not None else (ImagePool(0), ImagePool(0))
        random.seed(3)
        step = adversarial_seg_step_fused if fused else adversarial_seg_step
        for it, (bg, bn) in enumerate(_batches(B, N, 3)):
            torch.manual_seed(50 + it)
            l = step(g, d, gan, ce, opt, optD, tuple(t.to(DEV) for t in bg), tuple(t.to(DEV) for t in bn),
                     targs, history_pool_gt=pools[0], history_pool_nogt=pools[1])
            assert all(torch.isfinite(x) for x in l)
            if it == 0:
                grads[arm] = {k: v.grad.clone() for k, v in d.named_parameters()}
    # iteration 0: the pool is still filling, so it returns the inputs themselves (as fresh
    # leaves): the discriminator gradients equal the pool-free run's
    tol = 2e-5 if mode == "fp32" else 2e-3
    for k in grads[None]:
        assert rel_err(grads[fused][k], grads[None][k]) < tol, k


def test_forward_cache_does_not_confuse_batches_inside_a_step_scope():
    """Two different batches through PointNetCls(feature_transform=True) and
    PointNetSeg_regulization inside ONE weight_cache scope: temporaries of the first pass are freed
    and their addresses reused by the second; every output must equal the one computed alone."""
    torch.manual_seed(2)
    cls_net = M.PointNetCls(40, True).to(DEV).eval()
    seg_net = build_seg(3, 11, regu=True).to(DEV)
    ba = inputs(4, 256, 1)
    bb = inputs(4, 256, 2)
    alone = []
    for pts, _, _, cls in (ba, bb):
        with torch.no_grad():
            alone.append((cls_net(pts.to(DEV))[0].clone(), seg_net(pts.to(DEV), cls.to(DEV))[0].clone()))
    with weight_cache(), torch.no_grad():
        for rep in range(3):
            for i, (pts, _, _, cls) in enumerate((ba, bb)):
                a = cls_net(pts.to(DEV))[0]
                b = seg_net(pts.to(DEV), cls.to(DEV))[0]
                assert torch.equal(a, alone[i][0]) and torch.equal(b, alone[i][1]), (rep, i)


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_fused_ce_head_ignores_labels_like_cross_entropy(mode):
    """nn.CrossEntropyLoss() skips rows labelled ignore_index (-100) and averages over the rest;
    the fused head does the same (loss and gradients)."""
    g = build_seg(11, 12).to(DEV)
    g.precision = Precision(mode)
    pts, _, seg, cls = inputs(2, 300, 21)
    seg = seg.clone()
    seg[0, ::3] = -100
    seg[1, 5:50] = -100
    pts, seg, cls = pts.to(DEV), seg.to(DEV), cls.to(DEV)
    pred, _ = g(pts, cls)
    l_ref = F.cross_entropy(pred, seg)
    g.zero_grad(); l_ref.backward()
    ref = {k: v.grad.clone() for k, v in g.named_parameters()}
    loss, probs, _ = g.forward_ce(pts, cls, seg)
    g.zero_grad(); loss.backward()
    tol = 1e-5 if mode == "fp32" else 1e-3
    assert abs(loss.item() - l_ref.item()) < 10 * tol
    for k, v in g.named_parameters():
        assert rel_err(v.grad, ref[k]) < max(20 * tol, 2e-3 if mode != "fp32" else 0), k


def test_prefetch_pipeline_delivers_the_batch_with_device_jitter():
    """The input pipeline of the graphed step: what ``prefetch`` +
    ``step_prefetched`` put into the graph's static buffers is the pinned host batch -- bit for bit
    without augmentation, and the oracle's jitter stream of it (dataset/modelNetData.py:80-91) with
    ``set_jitter``; labels and one-hots are never touched."""
    from oracle import jitter_oracle as JO
    B, N = 2, 256
    g, d = _models(N, "fp16")
    g.to(DEV); d.to(DEV)
    opt, optD = _adam(g, d, True)
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=1e-3)
    batches = _batches(B, N, 3)
    pinned = [tuple(tuple(t.pin_memory() for t in part) for part in b) for b in batches]
    first = tuple(tuple(t.to(DEV) for t in part) for part in batches[0])
    gstep = GraphedAdversarialSegStep(g, d, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), opt, optD,
                                      targs, first[0], first[1], warmup=1, fused=True)
    gstep.prefetch(*pinned[1])
    gstep.step_prefetched()
    torch.cuda.synchronize()
    for dst, src in zip(gstep.static_gt + gstep.static_nogt, batches[1][0] + batches[1][1]):
        assert torch.equal(dst.cpu(), src)
    gstep.set_jitter(sigma=0.01, clip=0.05, seed=77)
    gstep.prefetch(*pinned[2])
    losses = gstep.step_prefetched()
    torch.cuda.synchronize()
    assert torch.isfinite(losses).all()
    (pts, cls, seg), (pts2, cls2) = batches[2]
    want_gt = JO.jitter(pts.numpy(), 0.01, 0.05, seed=77, offset=0)
    want_nogt = JO.jitter(pts2.numpy(), 0.01, 0.05, seed=77, offset=(pts.numel() + 3) // 4)
    assert (gstep.static_gt[0].cpu() - torch.from_numpy(want_gt)).abs().max().item() < 2e-7
    assert (gstep.static_nogt[0].cpu() - torch.from_numpy(want_nogt)).abs().max().item() < 2e-7
    assert torch.equal(gstep.static_gt[1].cpu(), cls) and torch.equal(gstep.static_gt[2].cpu(), seg)
    assert torch.equal(gstep.static_nogt[1].cpu(), cls2)


def test_data_parallel_graphed_step_two_ranks():
    """2 ranks, NCCL all-reduce captured inside the graph (tests/dist_graph_check.py): replicas stay
    identical and equal the 1-rank global-batch step.  Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run by hand: gpurun --gpus 2; result under profiles/)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731",
                          os.path.join(ROOT, "tests", "dist_graph_check.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_GRAPH_CHECK OK" in out.stdout

"""Part-IoU evaluation.

CPU: the numpy oracle (oracle/metric_oracle.py) against outputs of the reference's own
``utils.metric.batch_get_iou`` (tests/golden/metric_golden.npz).
GPU: ``pcadv_part_counts`` / ``pcadv_part_iou`` through the package's ``utils.metric`` mirror and
``trainer.run_testing_seg`` -- bit-exact against the golden values and the oracle (integer counts,
float64 quotients and sums in the reference's order)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import metric_oracle as MO

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["a", "b"]


@pytest.fixture(scope="module")
def mgold():
    return np.load(os.path.join(HERE, "golden", "metric_golden.npz"))


def _case(mgold, name):
    pred = mgold[name + "_pred"].astype(np.int64)
    seg = mgold[name + "_seg"].astype(np.int64)
    cats = mgold[name + "_cls"].astype(np.int64)
    cls = np.zeros((len(cats), 1, 16), np.float32)
    cls[np.arange(len(cats)), 0, cats] = 1.0
    return pred, seg, cls, mgold[name + "_iou"], mgold[name + "_correct"]


def _logits_for(pred, seed, C=50):
    """fp32 B x C x N logits whose (unique) argmax over C is ``pred``."""
    rng = np.random.default_rng(seed)
    B, N = pred.shape
    logits = rng.standard_normal((B, C, N)).astype(np.float32)
    logits[np.arange(B)[:, None], pred, np.arange(N)[None, :]] = 8.0 + rng.random((B, N)).astype(np.float32)
    return logits


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_metric(mgold, name):
    pred, seg, cls, iou, correct = _case(mgold, name)
    got = np.asarray(MO.batch_get_iou(pred, seg, cls[:, 0, :]), np.float64)
    assert got.tobytes() == iou.tobytes()                      # bit-exact float64
    pred_seg, accu, ious, cats = MO.evaluate(_logits_for(pred, 7), seg, cls, seg.shape[1])
    assert (pred_seg == pred).all()
    assert accu == correct.sum() / float(seg.shape[1])
    assert np.asarray(ious).tobytes() == iou.tobytes()
    assert (cats == mgold[name + "_cls"]).all()


# ----------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("name", CASES)
def test_device_iou_matches_reference_golden(mgold, name):
    from adversarial_learning_on_pointclouds_b200.utils import metric as DM
    pred, seg, cls, iou, correct = _case(mgold, name)
    d = lambda a: torch.from_numpy(a).cuda()
    # the reference-shaped entry point: int64 predictions in, list of floats out
    got = DM.batch_get_iou(d(pred), d(seg), d(cls[:, 0, :]))
    assert np.asarray(got, np.float64).tobytes() == iou.tobytes()
    # straight from B x C x N logits stored point-major (the generator's output view)
    logits = d(_logits_for(pred, 11)).transpose(1, 2).contiguous().transpose(1, 2)
    assert logits.stride() == (50 * seg.shape[1], 1, 50)
    g_iou, g_correct, g_cat, g_pred = DM.part_iou_from_logits(logits, d(seg), d(cls), want_pred=True)
    assert g_iou.cpu().numpy().tobytes() == iou.tobytes()
    assert (g_correct.cpu().numpy() == correct).all()
    assert (g_cat.cpu().numpy() == mgold[name + "_cls"]).all()
    assert (g_pred.cpu().numpy() == pred).all()
    # a channel-major (reference eager layout) tensor works too
    g2 = DM.part_iou_from_logits(d(_logits_for(pred, 11)), d(seg), d(cls))[0]
    assert g2.cpu().numpy().tobytes() == iou.tobytes()
    one = DM.get_iou(d(seg[3]), d(pred[3]), int(mgold[name + "_cls"][3]))
    assert one == iou[3]


@gpu
@pytest.mark.parametrize("B,N,C", [(1, 1, 50), (3, 127, 50), (5, 129, 50), (2, 1000, 13), (7, 4096, 64),
                                   (300, 40, 50)])
def test_device_counts_vs_oracle_random(B, N, C):
    """Uniform random labels and logits with exact ties (small integers): first-maximum argmax,
    counts and IoU against numpy."""
    from adversarial_learning_on_pointclouds_b200 import ops
    from adversarial_learning_on_pointclouds_b200.utils import metric as DM
    rng = np.random.default_rng(B * 1000 + N)
    logits = rng.integers(-3, 4, (B, N, C)).astype(np.float32)                 # many ties
    seg = rng.integers(0, C, (B, N)).astype(np.int64)
    cats = rng.integers(0, 16, B)
    cls = np.zeros((B, 16), np.float32)
    cls[np.arange(B), cats] = 1.0
    counts, correct, pred = ops.part_counts(torch.from_numpy(seg).cuda(), logits=torch.from_numpy(logits).cuda(),
                                            want_pred=True)
    ref_pred = np.argmax(logits, axis=2)
    assert (pred.cpu().numpy() == ref_pred).all()
    cn = counts.cpu().numpy()
    for l in range(C):
        assert (cn[:, 0, l] == ((ref_pred == l) & (seg == l)).sum(1)).all()
        assert (cn[:, 1, l] == (ref_pred == l).sum(1)).all()
        assert (cn[:, 2, l] == (seg == l).sum(1)).all()
    assert (correct.cpu().numpy() == (ref_pred == seg).sum(1)).all()
    if C >= 50:
        iou, cat = ops.part_iou(counts, torch.from_numpy(cls).cuda(), DM._part_begin(torch.device("cuda", 0)))
        want = np.asarray(MO.batch_get_iou(ref_pred, seg, cls), np.float64)
        assert iou.cpu().numpy().tobytes() == want.tobytes()
        assert (cat.cpu().numpy() == cats).all()


@gpu
def test_device_counts_padded_rows_nan_and_empty():
    from adversarial_learning_on_pointclouds_b200 import ops
    B, N, C = 4, 300, 50
    g = torch.Generator().manual_seed(5)
    wide = torch.randn(B, N, 64, generator=g).cuda()
    wide[..., C:] = 100.0                                                      # padding columns must be ignored
    logits = wide[..., :C]                                                     # row stride 64
    logits[1, 7, 20] = float("nan")                                            # torch.max: NaN wins
    seg = torch.randint(0, C, (B, N), generator=g).cuda()
    counts, correct, pred = ops.part_counts(seg, logits=logits, want_pred=True)
    want = logits.max(2)[1]
    assert want[1, 7].item() == 20
    assert torch.equal(pred, want)
    assert int(counts[:, 1].sum()) == B * N and int(counts[:, 2].sum()) == B * N
    assert torch.equal(correct.long(), (want == seg).sum(1))
    # empty batch / no points: zero counters, no launch
    c0, k0, _ = ops.part_counts(torch.zeros((0, 16), dtype=torch.int64, device="cuda"),
                                logits=torch.zeros((0, 16, C), device="cuda"))
    assert c0.shape == (0, 3, C) and k0.shape == (0,)
    with pytest.raises(ValueError):
        ops.part_counts(seg)
    with pytest.raises(ValueError):
        ops.part_counts(seg.cpu(), logits=logits)


@gpu
def test_full_size_count_properties():
    """cfg5 size (256 clouds x 4096 points): every point lands in exactly one pred and one gt bin,
    the hits equal the sum of the intersections, and the result is deterministic."""
    from adversarial_learning_on_pointclouds_b200 import ops
    B, N, C = 256, 4096, 50
    g = torch.Generator(device="cuda").manual_seed(3)
    logits = torch.randn(B, N, C, device="cuda", generator=g)
    seg = torch.randint(0, C, (B, N), device="cuda", generator=g)
    counts, correct, _ = ops.part_counts(seg, logits=logits)
    assert torch.equal(counts[:, 1].sum(1), torch.full((B,), N, dtype=torch.int64, device="cuda"))
    assert torch.equal(counts[:, 2].sum(1), torch.full((B,), N, dtype=torch.int64, device="cuda"))
    assert torch.equal(counts[:, 0].sum(1), correct.long())
    assert torch.equal(correct.long(), (logits.argmax(2) == seg).sum(1))
    again = ops.part_counts(seg, logits=logits)[0]
    assert torch.equal(again, counts)


@gpu
def test_run_testing_seg_matches_oracle(mgold):
    """trainer.run_testing_seg on the CUDA generator against the oracle's evaluation of the same
    logits: IoU means bit-exact given the predictions, accuracy exact, loss within fp32 rounding."""
    import torch.nn as nn
    from adversarial_learning_on_pointclouds_b200 import models as M, ops, trainer
    from adversarial_learning_on_pointclouds_b200.utils import init_net
    from helpers import inputs
    torch.manual_seed(0)
    G = init_net(M.PointNetSeg(50), "cuda", "xavier")
    G.precision = ops.Precision("fp32")
    batches = []
    for seed in (1234, 4321, 99):
        pts, _, seg, cls = inputs(6, 700, seed)
        batches.append((pts, cls, seg))
    args = types.SimpleNamespace(device="cuda", input_pts=700, tensorboard=False)
    acc, loss, cat_iou, all_iou = trainer.run_testing_seg(batches, list(range(18)), G, nn.CrossEntropyLoss(),
                                                          None, 0, None, args)
    # oracle side: numpy evaluation of the logits the CUDA model produced
    ious, cats, tot_acc, tot_loss = [], [], 0.0, 0.0
    G.eval()
    for pts, cls, seg in batches:
        with torch.no_grad():
            pred, _ = G(pts.cuda(), cls.cuda())
            tot_loss += nn.functional.cross_entropy(pred, seg.cuda()).item()
        _, accu, iou, cat = MO.evaluate(pred.cpu().numpy(), seg.numpy(), cls.numpy(), 700)
        tot_acc += accu
        ious += iou
        cats += list(cat)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want_cat, want_all = MO.summarize(ious, cats)
    assert acc == tot_acc / 18.0
    assert abs(loss - tot_loss / 18.0) < 1e-6
    assert all_iou == want_all
    assert (np.isnan(cat_iou) and np.isnan(want_cat)) or cat_iou == want_cat

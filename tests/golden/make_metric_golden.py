#!/usr/bin/env python3
"""Generate tests/golden/metric_golden.npz by running the UNMODIFIED reference's
``utils.metric.batch_get_iou`` (imported read-only from /root/reference) on seeded inputs.

Run in the build container only (``python tests/golden/make_metric_golden.py``); the GPU box has no
/root/reference and only reads the committed file.  The inputs themselves are stored (they are
small), so nothing depends on a generator reproducing them.
"""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
PART_BEGIN = [0, 4, 6, 8, 12, 16, 19, 22, 24, 28, 30, 36, 38, 41, 44, 47, 50]


def make_inputs(B, N, seed):
    """Clouds of every category (twice, then random); labels mostly inside the category's parts,
    with some parts absent from gt, from pred or from both, and a few out-of-category predictions."""
    rng = np.random.default_rng(seed)
    cats = np.concatenate([np.arange(16), np.arange(16), rng.integers(0, 16, B - 32)])
    seg = np.empty((B, N), np.int64)
    pred = np.empty((B, N), np.int64)
    for b, c in enumerate(cats):
        lo, hi = PART_BEGIN[c], PART_BEGIN[c + 1]
        parts = np.arange(lo, hi)
        gt_parts = parts if b % 3 else parts[:max(1, len(parts) - 1)]         # drop a part from gt
        pr_parts = parts if b % 4 else parts[:max(1, len(parts) - 1)]         # ... and from pred
        seg[b] = rng.choice(gt_parts, N)
        agree = rng.random(N) < 0.7
        pred[b] = np.where(agree, seg[b], rng.choice(pr_parts, N))
        if b % 4 == 0:                                                        # keep the dropped part out
            pred[b][pred[b] == parts[-1]] = parts[0]
        if b % 5 == 0:                                                        # stray labels of other categories
            stray = rng.random(N) < 0.05
            pred[b][stray] = rng.integers(0, 50, stray.sum())
    cls = np.zeros((B, 1, 16), np.float32)
    cls[np.arange(B), 0, cats] = 1.0
    # logits whose argmax is `pred` (distinct values, so the argmax is unambiguous)
    logits = rng.standard_normal((B, 50, N)).astype(np.float32)
    logits[np.arange(B)[:, None], pred, np.arange(N)[None, :]] = 8.0 + rng.random((B, N)).astype(np.float32)
    return logits, pred, seg, cls


def main():
    sys.path.insert(0, REF)
    from utils import metric as ref_metric                                    # numpy only
    out = {}
    for name, (B, N, seed) in {"a": (48, 512, 1234), "b": (40, 2048, 4321)}.items():
        logits, pred, seg, cls = make_inputs(B, N, seed)
        assert (np.argmax(logits, axis=1) == pred).all()
        ious = ref_metric.batch_get_iou(batch_pred=pred, batch_seg=seg, batch_cls=cls[:, 0, :])
        out[name + "_pred"] = pred.astype(np.int8)
        out[name + "_seg"] = seg.astype(np.int8)
        out[name + "_cls"] = np.argmax(cls[:, 0, :], axis=1).astype(np.int8)
        out[name + "_seed"] = np.array([B, N, seed])
        out[name + "_iou"] = np.asarray(ious, np.float64)
        out[name + "_correct"] = (pred == seg).sum(axis=1).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "metric_golden.npz"), **out)
    print("wrote metric_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

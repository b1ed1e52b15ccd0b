#!/usr/bin/env python3
"""On-device comparator ("also time the reference eager on one B200").

Runs the ORACLE's adversarial step (oracle/steps.py: the reference's ATen call sequence) with its
tensors on cuda:0 -- i.e. stock PyTorch eager kernels (cuDNN / cuBLAS), what a user of the reference
gets on this GPU -- next to this repo's step on the same synthetic batch, and prints one JSON line
per arm.  Test infrastructure: lives under tests/ because it executes oracle/; not collected by
pytest and not part of bench.py.  /root/reference is not needed.

    python tests/eager_gpu_comparator.py [--clouds 64 256] [--points 4096] [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def eager_arm(B, N, steps, warmup, tf32, autocast):
    from adversarial_learning_on_pointclouds_b200 import models as M
    from adversarial_learning_on_pointclouds_b200.utils import init_net
    from oracle import steps as S
    from helpers import inputs
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
    gp = S.leaf_params({k: v.cuda() for k, v in g.state_dict().items()})
    dp = S.leaf_params({k: v.cuda() for k, v in d.state_dict().items()})
    opt = torch.optim.Adam(list(gp.values()), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(list(dp.values()), lr=1e-5, betas=(0.9, 0.999))
    pts, _, seg, cls = inputs(B, N, 1234)
    pts2, _, _, cls2 = inputs(B, N, 4321)
    bg, bn = (pts.cuda(), cls.cuda(), seg.cuda()), (pts2.cuda(), cls2.cuda())

    def step():
        opt.zero_grad(); optD.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            S.adversarial_seg_step(gp, dp, bg, bn)
        opt.step(); optD.step()

    torch.cuda.reset_peak_memory_stats()
    ms = timed(step, steps, warmup)
    return ms, torch.cuda.max_memory_allocated() / 2**30


def ours_arm(B, N, steps, warmup):
    import torch.nn as nn
    import types
    from adversarial_learning_on_pointclouds_b200 import models as M
    from adversarial_learning_on_pointclouds_b200.trainer import GraphedAdversarialSegStep
    from adversarial_learning_on_pointclouds_b200.utils import init_net
    from helpers import inputs
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cuda", "xavier")
    d = init_net(M.PointwiseDiscNet(N, 50), "cuda", "xavier")
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999), fused=True, capturable=True)
    optD = torch.optim.Adam(d.parameters(), lr=1e-5, betas=(0.9, 0.999), fused=True, capturable=True)
    pts, _, seg, cls = inputs(B, N, 1234)
    pts2, _, _, cls2 = inputs(B, N, 4321)
    bg, bn = (pts.cuda(), cls.cuda(), seg.cuda()), (pts2.cuda(), cls2.cuda())
    args = types.SimpleNamespace(device="cuda", lambda_seg=1.0, lambda_adv=1e-3)
    torch.cuda.reset_peak_memory_stats()
    gs = GraphedAdversarialSegStep(g, d, nn.BCEWithLogitsLoss(), nn.CrossEntropyLoss(), opt, optD, args, bg, bn,
                                   device_labels=True, fused=True)
    ms = timed(lambda: gs(), steps, warmup)
    return ms, torch.cuda.max_memory_allocated() / 2**30


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, nargs="+", default=[64, 256])
    ap.add_argument("--points", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    for B in a.clouds:
        arms = [("torch eager fp32 (TF32 off)", lambda: eager_arm(B, a.points, a.steps, a.warmup, False, False)),
                ("torch eager TF32", lambda: eager_arm(B, a.points, a.steps, a.warmup, True, False)),
                ("torch eager bf16 autocast", lambda: eager_arm(B, a.points, a.steps, a.warmup, True, True)),
                ("pcadv fp16 fused step, CUDA graph", lambda: ours_arm(B, a.points, a.steps, a.warmup))]
        for name, fn in arms:
            try:
                ms, gib = fn()
                line = {"arm": name, "clouds": "%d+%d" % (B, B), "points": a.points, "ms_per_step": round(ms, 3),
                        "clouds_per_s": round(2 * B / ms * 1e3, 1), "peak_GiB": round(gib, 2)}
            except torch.OutOfMemoryError as exc:
                line = {"arm": name, "clouds": "%d+%d" % (B, B), "points": a.points,
                        "error": "out of memory: " + str(exc)[:100]}
            torch.cuda.empty_cache()
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

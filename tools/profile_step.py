#!/usr/bin/env python3
"""Per-kernel time table of one adversarial step (torch.profiler / CUPTI): every
kernel on the stream, libpcadv's and the trainer-side torch ops alike.  Used to
decide what to optimise next; the numbers of record come from bench.py (CUDA
events) and ncu (profiles/)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from torch.profiler import profile, ProfilerActivity

import bench
from adversarial_learning_on_pointclouds_b200 import models as M, Precision
from adversarial_learning_on_pointclouds_b200.utils import init_net
from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--device-labels", action="store_true")
    ap.add_argument("--shapes", action="store_true", help="also list torch ops by input shape")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    Bg, Bn, N = bench.WORKLOADS[args.workload]
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier").to(dev)
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier").to(dev)
    g.precision = d.precision = Precision(args.precision)
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, fused=True)
    optD = torch.optim.Adam(d.parameters(), lr=1e-5, fused=True)
    targs = argparse.Namespace(device=dev, lambda_seg=1.0, lambda_adv=1e-3)
    gan, seg = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()
    bg, bn = bench.synthetic_batches(Bg, Bn, N, 0)
    bg, bn = tuple(t.to(dev) for t in bg), tuple(t.to(dev) for t in bn)

    def step():
        adversarial_seg_step(g, d, gan, seg, opt, optD, bg, bn, targs, device_labels=args.device_labels)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=args.shapes) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    if args.shapes:
        ops_rows = []
        for e in prof.key_averages(group_by_input_shape=True):
            t = getattr(e, "device_time_total", 0) or 0
            if e.key.startswith("aten::") and t > 0:
                ops_rows.append((t / args.steps / 1e3, e.count / args.steps, e.key, str(e.input_shapes)[:90]))
        ops_rows.sort(reverse=True)
        print("torch ops by device time (ms/step):")
        for ms, cnt, key, shp in ops_rows[:30]:
            print("%8.3f ms  %5.1f  %-28s %s" % (ms, cnt, key, shp))
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if e.device_type.name == "CUDA" or (t and e.key.startswith(("void", "_ZN", "Memcpy", "Memset"))):
            rows.append((t / args.steps / 1e3, e.count / args.steps, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print("total device time per step: %.3f ms over %d kernel kinds" % (tot, len(rows)))
    for ms, cnt, key in rows[:args.top]:
        print("%8.3f ms  %6.1f calls  %s" % (ms, cnt, key[:110]))


if __name__ == "__main__":
    main()

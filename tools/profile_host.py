#!/usr/bin/env python3
"""cProfile of the host side of eager adversarial steps (what the unmodified reference trainer
pays per iteration when it drives the modules without the CUDA-graph wrapper)."""
import argparse, cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from adversarial_learning_on_pointclouds_b200 import models as M, Precision
from adversarial_learning_on_pointclouds_b200.utils import init_net
from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step

dev = torch.device("cuda", 0)
Bg, Bn, N = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg5"]
torch.manual_seed(0)
g = init_net(M.PointNetSeg(50), "cpu", "xavier").to(dev)
d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier").to(dev)
opt = torch.optim.Adam(g.parameters(), lr=1e-4, fused=True)
optD = torch.optim.Adam(d.parameters(), lr=1e-5, fused=True)
targs = argparse.Namespace(device=dev, lambda_seg=1.0, lambda_adv=1e-3)
gan, seg = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()
bg, bn = bench.synthetic_batches(Bg, Bn, N, 0)
bg, bn = tuple(t.to(dev) for t in bg), tuple(t.to(dev) for t in bn)
step = lambda: adversarial_seg_step(g, d, gan, seg, opt, optD, bg, bn, targs, device_labels=True)
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t_issue = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
t_total = (time.perf_counter() - t0) / 5
print("host issue time per step %.2f ms, wall per step %.2f ms" % (t_issue * 1e3, t_total * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

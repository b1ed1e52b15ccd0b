#!/usr/bin/env python3
"""Diagnostic: where the gap between the resident and the end-to-end loop of bench.py comes from.
Times the graphed cfg5 step with pieces of the input pipeline switched on one at a time."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from adversarial_learning_on_pointclouds_b200 import models as M, Precision
from adversarial_learning_on_pointclouds_b200.utils import init_net
from adversarial_learning_on_pointclouds_b200.trainer import GraphedAdversarialSegStep

dev = torch.device("cuda", 0)
Bg, Bn, N = 256, 256, 4096
torch.manual_seed(0)
g = init_net(M.PointNetSeg(50), "cpu", "xavier").to(dev)
d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier").to(dev)
g.precision = d.precision = Precision("fp16")
opt = torch.optim.Adam(g.parameters(), lr=1e-4, fused=True, capturable=True)
optD = torch.optim.Adam(d.parameters(), lr=1e-5, fused=True, capturable=True)
targs = argparse.Namespace(device=dev, lambda_seg=1.0, lambda_adv=1e-3)
hg, hn = bench.synthetic_batches(Bg, Bn, N, 0)
hg, hn = tuple(t.pin_memory() for t in hg), tuple(t.pin_memory() for t in hn)
dg, dn = tuple(t.to(dev) for t in hg), tuple(t.to(dev) for t in hn)
device_labels = "--device-labels" in sys.argv
gs = GraphedAdversarialSegStep(g, d, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), opt, optD, targs, dg, dn,
                               warmup=2, fused=True, device_labels=device_labels)
lh = [torch.empty(3).pin_memory() for _ in range(2)]
ev = [torch.cuda.Event(), torch.cuda.Event()]


def timed(fn, n=30):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (t1 - t0) / n * 1e3


def replay_only(i):
    gs.graph.replay()

def resident(i):
    gs()

def with_prefetch(i):
    gs.step_prefetched(); gs.prefetch(hg, hn)

def with_read_sync(i):
    l = gs.step_prefetched(); gs.prefetch(hg, hn)
    lh[0].copy_(l, non_blocking=True); torch.cuda.current_stream().synchronize()

def with_read_lag(i):
    l = gs.step_prefetched(); gs.prefetch(hg, hn)
    lh[i & 1].copy_(l, non_blocking=True); ev[i & 1].record()
    if i > 0:
        ev[(i - 1) & 1].synchronize()

gs.prefetch(hg, hn)
for name, fn in (("graph.replay() only", replay_only), ("resident gstep()", resident), ("+ prefetch / step_prefetched", with_prefetch),
                 ("+ loss read, sync every step", with_read_sync), ("+ loss read, lag 1", with_read_lag), ("resident again", resident)):
    gpu_ms, host_ms = timed(fn)
    print("%-34s GPU %.3f ms/step   host enqueue %.3f ms/step" % (name, gpu_ms, host_ms))

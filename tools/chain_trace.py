#!/usr/bin/env python3
"""Cycle-stamped timeline of tc_chain_kernel's CTA 0 (debug hook pcadv_debug_chain_trace): where the
serial MMA -> epilogue -> MMA chain of a tile spends its time.  Prints, per (tile round, layer), the
MMA thread's wait / issue stamps and the epilogue warps' wake / TMEM-load / math / arrive stamps,
relative to the first stamp."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from adversarial_learning_on_pointclouds_b200 import _lib, ops
from adversarial_learning_on_pointclouds_b200.ops import ACT_RELU

P, k0, widths = 1 << 20, 64, [128, 128, 128]
dev = "cuda"
x = (torch.randn((P, k0), device=dev)).half()
layers, k = [], k0
for n in widths:
    layers.append(((torch.randn((n, k), device=dev) * 0.05).half(), torch.randn(n, device=dev), ACT_RELU, 0.0))
    k = n
for _ in range(3):
    ops.chain(x, layers)
torch.cuda.synchronize()
SLOTS = 256
trace = torch.zeros((10, SLOTS), dtype=torch.int64, device=dev)
fn = _lib.lib().pcadv_debug_chain_trace
fn.argtypes, fn.restype = [C.c_void_p], None
fn(C.c_void_p(trace.data_ptr()))
ops.chain(x, layers)
torch.cuda.synchronize()
fn(C.c_void_p(0))
t = trace.cpu().numpy()
t0 = t[t > 0].min()
NL = len(widths)
rel = lambda v: int(v - t0) if v > 0 else -1
print("MMA thread (warp 1): per (round, layer, half): [wait start, in_ready wake, MMAs issued + commit]")
for r in range(3, 7):
    for l in range(NL):
        for h in range(2):
            s = ((r * NL + l) * 2 + h) * 3
            if s + 2 < SLOTS:
                print("  round %d layer %d half %d: %s" % (r, l, h, [rel(t[1, s + j]) for j in range(3)]))
print("epilogue warps: per (round, layer): [wait start, acc_full wake, first TMEM loads done, first step packed,"
      " layer done, arrived]")
for w in (2, 6):
    for r in range(3, 7):
        for l in range(NL):
            s = (r * NL + l) * 6
            if s + 5 < SLOTS:
                v = [rel(t[w, s + j]) for j in range(6)]
                d = [v[j + 1] - v[j] for j in range(5)]
                print("  warp %d (half %d) round %d layer %d: %s  deltas %s" % (w, (w - 2) >> 2, r, l, v, d))

#!/usr/bin/env python3
"""Forward + backward timing of the other BASELINE.json configs on one GPU (eager launches, CUDA
events): cfg1 PointNetCls(40) B=32 N=2500, cfg2 PointNetDenseCls(50) B=32 N=2500, cfg4
PointNetCls(40, feature_transform=True) + regulariser B=128 N=2048.  Not the headline bench
(bench.py measures cfg5); used to spot slow paths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from adversarial_learning_on_pointclouds_b200 import models as M, ops, Precision
import bench

dev = "cuda"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    prec = Precision(sys.argv[1] if len(sys.argv) > 1 else "fp16")
    torch.manual_seed(0)
    rows = []
    # cfg1
    B, N = 32, 2500
    pts, y, seg, cls = bench.synthetic_inputs(B, N, 1234)
    m = M.PointNetCls(40, False).to(dev); m.precision = prec
    P, Y = pts.to(dev), y.to(dev)
    def f1():
        m.zero_grad(set_to_none=True)
        logits, _, _ = m(P)
        F.cross_entropy(logits, Y).backward()
    rows.append(("cfg1 PointNetCls B=32 N=2500", B, timeit(f1)))
    # cfg2
    m2 = M.PointNetDenseCls(50, False).to(dev); m2.precision = prec
    S = seg.to(dev)
    X = P.transpose(1, 2).contiguous()
    def f2():
        m2.zero_grad(set_to_none=True)
        lp, _ = m2(X)
        F.nll_loss(lp.reshape(-1, 50), S.reshape(-1)).backward()
    rows.append(("cfg2 PointNetDenseCls B=32 N=2500", B, timeit(f2)))
    # cfg4
    B4, N4 = 128, 2048
    pts4, y4, _, _ = bench.synthetic_inputs(B4, N4, 1234)
    m4 = M.PointNetCls(40, True).to(dev); m4.precision = prec
    P4, Y4 = pts4.to(dev), y4.to(dev)
    def f4():
        m4.zero_grad(set_to_none=True)
        logits, _, tf = m4(P4)
        (F.cross_entropy(logits, Y4) + 1e-3 * M.feature_transform_regularizer(tf)).backward()
    rows.append(("cfg4 PointNetCls(ft) B=128 N=2048", B4, timeit(f4)))
    for name, b, ms in rows:
        print("%-36s %8.3f ms/step  %10.0f clouds/s" % (name, ms, b / ms * 1e3))
    if "--kernels" in sys.argv:
        with ops.KernelTimer() as kt:
            f4()
        for tag, (calls, ms, _rows) in sorted(kt.summary().items(), key=lambda kv: -kv[1][1])[:25]:
            print("   %-40s %3d calls %8.3f ms" % (tag, calls, ms))


if __name__ == "__main__":
    main()

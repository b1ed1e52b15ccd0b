#!/usr/bin/env python3
"""Same-process A/B of the generator's trunk backward (levels 3, 2, 1 + fc1's weight gradient) issued as
separate dgrad / wgrad launches versus pcadv_backlevel launches, back to back (no host gaps between
the kernels of a sequence), alternating the two so that thermal drift hits both alike."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from adversarial_learning_on_pointclouds_b200 import ops
from adversarial_learning_on_pointclouds_b200.ops import ACT_RELU, ENGINE_TC

DEV = "cuda"


def r16(shape, scale=1.0):
    return (torch.randn(shape, device=DEV) * scale).half()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 21)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--only-head", action="store_true", help="the single-segment levels only (ncu captures)")
    args = ap.parse_args()
    P, N = args.points, 4096
    torch.manual_seed(0)
    xs = [r16((P, k)).relu_() for k in (64, 128, 128, 128, 512)]
    bits = [torch.randint(-2 ** 31, 2 ** 31 - 1, (P, k // 32), device=DEV, dtype=torch.int32) for k in (64, 128, 128)]
    dz4 = r16((P, 128))
    dh1 = r16((P, 256))
    wts = [r16((n, 384), 0.05) for n in (64, 128, 128)]           # dgrad weights of levels 1, 2, 3
    dw1 = torch.zeros((256, 3024), device=DEV)
    dws = [torch.zeros((128, n), device=DEV) for n in (64, 128, 128)]
    dbs = [torch.zeros((128,), device=DEV) for _ in range(3)]
    dcb = torch.zeros((P // N, 256), device=DEV)
    s1 = torch.ones(1, device=DEV)
    sl = ((0, 64), (64, 192), (192, 320))

    def old():
        ops.wgrad(dh1, xs, dw=dw1[:, :960], dgroup_bias=dcb, rows_per_group=N, scale=s1, engine=ENGINE_TC)
        dz = dz4
        for li in (2, 1, 0):
            ops.wgrad(dz, [xs[li]], dw=dws[li], dbias=dbs[li], scale=s1, engine=ENGINE_TC)
            dz, _, _ = ops.linear([dz, dh1], wts[li], mask=xs[li], mask_act=ACT_RELU, out_dtype=torch.float16,
                                  engine=ENGINE_TC, mask_bits=bits[li])
        return dz

    def new():
        ops.wgrad(dh1, xs[3:], dw=dw1[:, 320:960], dgroup_bias=dcb, rows_per_group=N, scale=s1, engine=ENGINE_TC)
        dz = dz4
        for li in (2, 1, 0):
            dz = ops.backlevel([dz, dh1], wts[li], xs[li], mask_bits=bits[li], dws=[dws[li], dw1[:, sl[li][0]:sl[li][1]]],
                               dbiases=[dbs[li], None], scale=s1)
        return dz

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    def pair(name, k, n, rows):
        """one single-segment level: dz [rows, k] -> dz_out [rows, n]"""
        dz = r16((rows, k))
        x = r16((rows, n)).relu_()
        bt = torch.randint(-2 ** 31, 2 ** 31 - 1, (rows, n // 32), device=DEV, dtype=torch.int32)
        wt = r16((n, k), 0.05)
        dw = torch.zeros((k, n), device=DEV)
        db = torch.zeros((k,), device=DEV)

        def o():
            ops.wgrad(dz, [x], dw=dw, dbias=db, scale=s1, engine=ENGINE_TC)
            ops.linear([dz], wt, mask=x, mask_act=ACT_RELU, out_dtype=torch.float16, engine=ENGINE_TC, mask_bits=bt)

        def nw():
            ops.backlevel([dz], wt, x, mask_bits=bt, dws=[dw], dbiases=[db], scale=s1)
        return name, o, nw

    def rowmax_case(rows):
        """backward of the discriminator's conv4 (64 -> 128) + ReLU + max over channels and the dgrad to conv3"""
        n, k = 128, 64
        dy = torch.randn(rows, device=DEV)
        val = torch.rand(rows, device=DEV)
        idx = torch.randint(0, n, (rows,), device=DEV, dtype=torch.int32)
        y = r16((rows, k)).relu_()
        bt = torch.randint(-2 ** 31, 2 ** 31 - 1, (rows, k // 32), device=DEV, dtype=torch.int32)
        w = r16((n, k), 0.1)
        wt = w.t().contiguous()
        dw = torch.zeros((n, k), device=DEV)
        db = torch.zeros((n,), device=DEV)

        def o():
            ops.rowmax_wgrad(dy, val, idx, y, n, act=ACT_RELU, dw=dw, dbias=db)
            ops.rowmax_dgrad(dy, val, idx, w, y, act=ACT_RELU, scale=s1, prev_act=ACT_RELU, out_dtype=torch.float16)

        def nw():
            dz = ops.rowmax_bwd(dy, val, idx, n, act=ACT_RELU, scale=s1, out_dtype=torch.float16)
            ops.backlevel([dz], wt, y, mask_bits=bt, dws=[dw], dbiases=[db], scale=s1)
        def nw2():
            ops.backlevel(None, wt, y, mask_bits=bt, dws=[dw], dbiases=[db], scale=s1,
                          onehot=(dy, val, idx, n, ACT_RELU, 0.0, s1))
        if os.environ.get("LEVEL_AB_ONEHOT", "1") != "0":
            return "rowmax level n128 k64 (gather | one-hot in kernel)", o, nw2
        return "rowmax level n128 k64 (gather | one-hot)", o, nw

    cases = [("trunk levels 3,2,1 + fc1 wgrad", old, new), rowmax_case(P // 2), pair("fc4 level k64 n128", 64, 128, P),
             pair("fc3 level k128 n256", 128, 256, P), pair("fc2 level k256 n256", 256, 256, P),
             pair("disc level k64 n64", 64, 64, P // 2)]
    if args.only_head:
        cases = cases[1:]
    for name, o, nw in cases:
        o(); nw()
        torch.cuda.synchronize()
        for r in range(args.rounds):
            a = timed(o)
            b = timed(nw)
            print("%-32s round %d: separate %.3f ms   backlevel %.3f ms" % (name, r, a, b))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Micro-benchmark of single libpcadv ops at the cfg5 shapes (rows = 2^20 points),
timed with CUDA events on the launching stream after warm-up.  Prints achieved
GB/s (algorithmic bytes) and TFLOP/s per op.  ``--only NAME`` runs one case (for
an ncu capture: ``ncu --set full -k regex:tc_linear ... python tools/microbench.py
--only conv5 --iters 1``)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from adversarial_learning_on_pointclouds_b200 import ops
from adversarial_learning_on_pointclouds_b200.ops import ACT_NONE, ACT_RELU, ENGINE_TC, ENGINE_SIMT

DEV = "cuda"


def r16(shape, scale=1.0):
    return (torch.randn(shape, device=DEV) * scale).half()


def cases(P, N):
    B = P // N
    c = {}

    def lin(name, ks, n, out_dtype=torch.float16, mask=False, addend=False, bias=True, colmax=False,
            rowmax=False, want_out=True, gb=False, bits=False, maskbits=False):
        segs = [r16((P, k)) for k in ks]
        w = r16((n, sum(ks)), 0.05)
        kw = dict(bias=torch.randn(n, device=DEV) if bias else None,
                  act=ACT_NONE if (maskbits or mask) else ACT_RELU, out_dtype=out_dtype,
                  engine=ENGINE_TC, colmax=colmax, rowmax=rowmax, want_out=want_out,
                  rows_per_group=N if (colmax or gb) else 0)
        if mask:
            kw.update(mask=r16((P, n)), mask_act=ACT_RELU)
        if addend:
            kw.update(addend=torch.randn((P, n), device=DEV))
        if gb:
            kw.update(group_bias=torch.randn((B, n), device=DEV))
        if bits:
            kw.update(bits_out=ops.new_bits(P, n, DEV))
        if maskbits:
            kw.update(mask=r16((P, n)), mask_act=ACT_RELU,
                      mask_bits=torch.randint(-2 ** 31, 2 ** 31 - 1, (P, n // 32), device=DEV, dtype=torch.int32))
        esz = 2 if out_dtype == torch.float16 else 4
        nbytes = P * (2 * sum(ks) + (esz * n if want_out else 0) + (2 * n if mask else 0) +
                      (4 * n if addend else 0))
        flops = 2.0 * P * sum(ks) * n
        c[name] = (lambda: ops.linear(segs, w, **kw), nbytes, flops)

    lin("conv2", [64], 128)
    lin("conv3", [128], 128)
    lin("conv5", [128], 512)
    lin("conv5_bits", [128], 512, bits=True)
    lin("conv3_bits", [128], 128, bits=True)
    lin("fc2_bits", [256], 256, bits=True)
    lin("dz_fc1_mb", [256], 256, maskbits=True, bias=False)
    lin("dz4_mb", [512, 256], 128, maskbits=True, bias=False)
    lin("dz5_mb", [256], 512, maskbits=True, bias=False)
    lin("d2_bits", [64], 64, bits=True)
    lin("conv6max", [512], 2048, colmax=True, want_out=False)
    lin("fc1", [64, 128, 128, 128, 512], 256, gb=True, bias=False)
    lin("fc2", [256], 256)
    lin("fc3", [256], 128)
    lin("fc4", [128], 50, out_dtype=torch.float32)
    lin("dz_fc3", [64], 128, mask=True, bias=False)
    lin("dz_fc2", [128], 256, mask=True, bias=False)
    lin("dz_fc1", [256], 256, mask=True, bias=False)
    lin("dz5", [256], 512, mask=True, addend=True, bias=False)
    lin("dz4", [512, 256], 128, mask=True, bias=False)
    lin("dz3", [128, 256], 128, mask=True, bias=False)
    lin("dz1", [128, 256], 64, mask=True, bias=False)
    lin("disc4max", [64], 128, rowmax=True, want_out=False)

    # chained narrow layers (pcadv_chain): the generator trunk conv2 -> conv4 and the head tail
    def chain(name, k0, widths, last_f32=False, rowmax=False):
        x = r16((P, k0))
        layers, k = [], k0
        for i, n in enumerate(widths):
            act = ACT_NONE if (last_f32 and i == len(widths) - 1) else ACT_RELU
            layers.append((r16((n, k), 0.05), torch.randn(n, device=DEV), act, 0.0))
            k = n
        nb = P * 2 * k0 + sum(P * (4 * n if (last_f32 and i == len(widths) - 1) else 2 * n + n // 8)
                              for i, n in enumerate(widths))
        fl = 2.0 * P * sum(a * b for a, b in zip([k0] + widths[:-1], widths))
        if rowmax:      # the last layer stores nothing but the 8-byte key
            nb -= P * (2 * widths[-1] + widths[-1] // 8) - 8 * P
        c[name] = (lambda: ops.chain(x, layers, last_f32=last_f32, rowmax=rowmax), nb, fl)

    chain("chain_trunk", 64, [128, 128, 128])
    chain("chain_tail", 256, [128, 50], last_f32=True)
    chain("chain_disc", 64, [64, 64, 64, 128], rowmax=True)

    def wg(name, n, ks, dbias=True):
        dz = r16((P, n))
        segs = [r16((P, k)) for k in ks]
        dw = torch.zeros((n, sum(ks)), device=DEV)
        db = torch.zeros((n,), device=DEV) if dbias else None
        nbytes = P * 2 * (n + sum(ks))
        c[name] = (lambda: ops.wgrad(dz, segs, dw=dw, dbias=db, engine=ENGINE_TC), nbytes,
                   2.0 * P * n * sum(ks))

    wg("wg_fc1", 256, [64, 128, 128, 128, 512], dbias=False)
    dzg = r16((P, 256))
    segg = [r16((P, k)) for k in [64, 128, 128, 128, 512]]
    dwg = torch.zeros((256, 3024), device=DEV)
    dcb = torch.zeros((B, 256), device=DEV)
    s2 = torch.ones(2, device=DEV)
    c["wg_fc1_g"] = (lambda: ops.wgrad(dzg, segg, dw=dwg[:, :960], dgroup_bias=dcb, rows_per_group=N,
                                       scale=s2[1:2], engine=ENGINE_TC), P * 2 * (256 + 960), 2.0 * P * 256 * 960)
    wg("wg_conv5", 512, [128])
    wg("wg_conv5_swapped", 128, [512], dbias=False)      # x4 as the M side, dz5 as eight N boxes: every byte once
    wg("wg_fc2_swapped", 256, [256], dbias=False)
    wg("wg_fc2", 256, [256])
    wg("wg_conv3", 128, [128])
    wg("wg_d2", 64, [64])

    # fused backward levels (pcadv_backlevel): dgrad of the level + weight gradients of x's consumers
    def lv(name, ks, n, group=False, nsum=1):
        segs = [r16((P, k)) for k in ks]
        w = r16((n, sum(ks)), 0.05)
        x = r16((P, n)).relu_()
        bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (P, n // 32), device=DEV, dtype=torch.int32)
        dws = [torch.zeros((k, n), device=DEV) for k in ks]
        dbs = [torch.zeros((k,), device=DEV) if i < nsum else None for i, k in enumerate(ks)]
        dgs = [torch.zeros((B, k), device=DEV) if (group and i == 0) else None for i, k in enumerate(ks)]
        s1 = torch.ones(1, device=DEV)
        out = torch.empty((P, n), device=DEV, dtype=torch.float16)
        nbytes = P * (2 * sum(ks) + 4 * n + n // 8)
        c[name] = (lambda: ops.backlevel(segs, w, x, mask_bits=bits, dws=dws, dbiases=dbs, dgroups=dgs,
                                         rows_per_group=N, scale=s1, out=out), nbytes, 4.0 * P * sum(ks) * n)

    lv("lv_fc4", [64], 128)
    lv("lv_d", [64], 64)
    lv("lv_d_nosum", [64], 64, nsum=0)
    lv("lv_fc3", [128], 256)
    lv("lv_fc2", [256], 256)
    lv("lv5", [256], 512, group=True, nsum=0)
    lv("lv4", [512, 256], 128)
    lv("lv3", [128, 256], 128)
    lv("lv4_nosum", [512, 256], 128, nsum=0)
    lv("lv3_nosum", [128, 256], 128, nsum=0)
    lv("lv1", [128, 256], 64)

    # sparse max-pool backward (conv6): realistic argmax distribution from a real forward
    x5 = r16((P, 512)).relu_()
    w6 = r16((2048, 512), 0.05)
    _, key, _ = ops.linear([x5], w6, want_out=False, colmax=True, rows_per_group=N, engine=ENGINE_TC)
    g6, idx6 = ops.max_finalize(key, ACT_RELU)
    # realistic argmax distribution (measured on the reference at xavier init, N = 4096: ~430
    # touched rows per cloud, the hottest row the argmax of 100-150 channels)
    pr = 1.0 / (torch.arange(432, device=DEV, dtype=torch.float32) + 6.0)
    pick = torch.multinomial(pr.expand(B, -1), 2048, replacement=True)
    rows_sel = torch.stack([torch.randperm(N, device=DEV)[:432] for _ in range(B)])
    idx6 = torch.gather(rows_sel, 1, pick).to(torch.int32).contiguous()
    g6 = g6.clamp_min(1e-3)
    dg6 = torch.randn((B, 2048), device=DEV)
    dw6 = torch.zeros((2048, 512), device=DEV)
    db6 = torch.zeros((2048,), device=DEV)
    dz5 = r16((P, 512))
    c["maxbwd_dw"] = (lambda: ops.maxpool_bwd(dg6, g6, idx6, x5, w6, N, act=ACT_RELU, dw=dw6, dbias=db6),
                      B * 2048 * 1024, 2.0 * B * 2048 * 512)
    c["maxbwd_rows"] = (lambda: ops.maxpool_bwd(dg6, g6, idx6, x5, w6, N, act=ACT_RELU, dz_inout=dz5,
                                                prev_act=ACT_RELU), B * 2048 * 1024, 2.0 * B * 2048 * 512)

    # evaluation: argmax over 50 logits + per-cloud part counts + IoU (utils/metric.py on the device)
    from adversarial_learning_on_pointclouds_b200.utils import metric as DM
    lg = torch.randn((B, N, 50), device=DEV)
    sg = torch.randint(0, 50, (B, N), device=DEV)
    oh = torch.zeros((B, 16), device=DEV)
    oh[torch.arange(B), torch.randint(0, 16, (B,), device=DEV)] = 1.0
    c["part_iou"] = (lambda: DM.part_iou_from_logits(lg, sg, oh), P * (200 + 8), 0.0)

    x50 = torch.randn((P, 50), device=DEV)
    w50 = torch.randn((64, 50), device=DEV) * 0.1
    c["disc1_simt"] = (lambda: ops.linear([x50], w50, bias=torch.zeros(64, device=DEV), act=ACT_RELU,
                                          out_dtype=torch.float16, engine=ENGINE_SIMT),
                       P * (200 + 128), 2.0 * P * 50 * 64)
    pts = torch.randn((P, 3), device=DEV)
    w3 = torch.randn((64, 3), device=DEV)
    c["conv1_simt"] = (lambda: ops.linear([pts], w3, bias=torch.zeros(64, device=DEV), act=ACT_RELU,
                                          out_dtype=torch.float16, engine=ENGINE_SIMT),
                       P * (12 + 128), 2.0 * P * 3 * 64)
    s2 = torch.ones(2, device=DEV)
    dy = torch.randn(P, device=DEV)
    val = torch.rand(P, device=DEV)
    idx = torch.randint(0, 128, (P,), device=DEV, dtype=torch.int32)
    c["rowmax_bwd"] = (lambda: ops.rowmax_bwd(dy, val, idx, 128, act=ACT_RELU, out_dtype=torch.float16),
                       P * (12 + 256), 0.0)
    # fused loss heads over the fp32 logits (utils/trainer.py:899-901, :914)
    lg50 = torch.randn((P, 50), device=DEV) * 0.3
    lab = torch.randint(0, 50, (P,), device=DEV)
    acc2 = torch.zeros(2, device=DEV)
    pr16, dz16 = torch.empty((P, 64), device=DEV, dtype=torch.float16), torch.empty((P, 64), device=DEV, dtype=torch.float16)
    c["head_ce"] = (lambda: ops.softmax_head(lg50, ops.HEAD_CE, labels=lab, out_dtype=torch.float16, cols=64,
                                             dz_gain=256.0, loss_sum=acc2[0:1], valid_count=acc2[1:2],
                                             probs_out=pr16, dz_out=dz16), P * (200 + 8 + 256), 0.0)
    c["head_lsm"] = (lambda: ops.softmax_head(lg50, ops.HEAD_LSM, out_dtype=torch.float16, cols=64, probs_out=pr16),
                     P * (200 + 128), 0.0)
    dy16 = r16((P, 64))
    c["lsm_bwd"] = (lambda: ops.logsoftmax_bwd(pr16, dy16, 50, cols=64, out=dz16), P * 384, 0.0)
    # discriminator: backward of conv4 + ReLU + max over channels (gather kernels)
    yprev = r16((P, 64)).relu_()
    wd4 = r16((128, 64), 0.1)
    dwd = torch.zeros((128, 64), device=DEV)
    dbd = torch.zeros((128,), device=DEV)
    c["rowmax_wgrad"] = (lambda: ops.rowmax_wgrad(dy, val, idx, yprev, 128, act=ACT_RELU, dw=dwd, dbias=dbd),
                         P * (12 + 128), 2.0 * P * 64)
    c["rowmax_dgrad"] = (lambda: ops.rowmax_dgrad(dy, val, idx, wd4, yprev, act=ACT_RELU, scale=s2[0:1],
                                                  prev_act=ACT_RELU, out_dtype=torch.float16),
                         P * (12 + 128 + 128), 2.0 * P * 64)
    dz64 = r16((P, 64))
    dw3 = torch.zeros((64, 3), device=DEV)
    db3 = torch.zeros((64,), device=DEV)
    c["conv1_wgrad"] = (lambda: ops.wgrad(dz64, [pts], dw=dw3, dbias=db3, scale=s2[1:2], engine=ENGINE_SIMT),
                        P * (12 + 128), 2.0 * P * 3 * 64)
    dl = torch.randn((P, 50), device=DEV) * 1e-6
    c["amax"] = (lambda: ops.amax_scale(dl), P * 200, 0.0)
    s2 = torch.ones(2, device=DEV)
    c["convert50"] = (lambda: ops.convert(dl, torch.float16, cols_pad=64, scale=s2[0:1]), P * 328, 0.0)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 20)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    torch.manual_seed(0)
    cs = cases(args.points, args.n)
    print("%-12s %9s %9s %9s" % ("op", "ms", "GB/s", "TFLOP/s"))
    for name, (fn, nbytes, flops) in cs.items():
        if args.only and name not in args.only.split(","):
            continue
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        print("%-12s %9.3f %9.0f %9.1f" % (name, ms, nbytes / ms / 1e6, flops / ms / 1e9))


if __name__ == "__main__":
    main()

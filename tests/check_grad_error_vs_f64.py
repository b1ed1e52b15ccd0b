#!/usr/bin/env python3
"""Diagnostic (not collected by pytest): per-parameter gradient error of one adversarial step against
a float64 evaluation of the oracle, for (a) the fp32 oracle on the CPU, (b) the fp32 oracle run by
stock PyTorch eager on the GPU (TF32 off), (c) the same with TF32 tensor cores (cuDNN / cuBLAS
defaults of the speed comparison), (d) the same under bf16 and (e) fp16 autocast, (f) this repo's fp32
verification mode, (g) its fp16 mode through the reference-shaped loop body and (h) through the fused
one-pass step that bench.py times, (i) its bf16 mode.  (c)-(e) are the yardstick for the 16-bit modes:
what the stock tensor-core paths do to the same un-conditioned gradients.
Test infrastructure: executes oracle/.

    python tests/check_grad_error_vs_f64.py [--clouds 3] [--points 384]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=3)
    ap.add_argument("--points", type=int, default=384)
    a = ap.parse_args()
    from adversarial_learning_on_pointclouds_b200 import models as M
    from adversarial_learning_on_pointclouds_b200.ops import Precision
    from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step
    from adversarial_learning_on_pointclouds_b200.utils import init_net
    from oracle import steps
    from helpers import inputs, randomize_biases
    import types
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    B, N = a.clouds, a.points
    torch.manual_seed(5)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
    randomize_biases([g, d], 4)
    pts, _, seg, cls = inputs(B, N, 100)
    pts2, _, _, cls2 = inputs(B, N, 500)
    lab = (torch.empty(B, N).uniform_(0.7, 1.05), torch.empty(B, N).uniform_(0.0, 0.305))

    def oracle(dev, dtype, tf32=False, autocast=None):
        import contextlib
        cast = lambda t: t.to(dev, dtype) if t.is_floating_point() else t.to(dev)
        gp = steps.leaf_params({k: cast(v) for k, v in g.state_dict().items()})
        dp = steps.leaf_params({k: cast(v) for k, v in d.state_dict().items()})
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
        ctx = torch.autocast("cuda", dtype=autocast) if autocast is not None else contextlib.nullcontext()
        try:
            with ctx:
                steps.adversarial_seg_step(gp, dp, tuple(cast(t) for t in (pts, cls, seg)),
                                           tuple(cast(t) for t in (pts2, cls2)),
                                           labels=tuple(cast(t) for t in lab))
        finally:
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        out = {"G." + k: v.grad.double().cpu() for k, v in gp.items()}
        out.update({"D." + k: v.grad.double().cpu() for k, v in dp.items()})
        return out

    def ours(mode, fused=False):
        import copy
        from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step_fused
        gg, dd = copy.deepcopy(g).cuda(), copy.deepcopy(d).cuda()
        gg.precision = dd.precision = Precision(mode)
        opt = torch.optim.SGD(gg.parameters(), lr=0.0)
        optD = torch.optim.SGD(dd.parameters(), lr=0.0)
        targs = types.SimpleNamespace(device="cuda", lambda_seg=1.0, lambda_adv=1e-3)
        labels = iter([None, lab[0].cuda(), lab[1].cuda()])

        def label_fn(d_out, value, random):
            nxt = next(labels)
            return torch.full_like(d_out, float(value)) if not random else nxt

        step = adversarial_seg_step_fused if fused else adversarial_seg_step
        step(gg, dd, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), opt, optD,
             tuple(t.cuda() for t in (pts, cls, seg)), tuple(t.cuda() for t in (pts2, cls2)),
             targs, label_fn=label_fn)
        out = {"G." + k: v.grad.double().cpu() for k, v in gg.named_parameters()}
        out.update({"D." + k: v.grad.double().cpu() for k, v in dd.named_parameters()})
        return out

    truth = oracle("cpu", torch.float64)
    arms = {"cpu fp32": oracle("cpu", torch.float32), "eager fp32": oracle("cuda", torch.float32),
            "eager tf32": oracle("cuda", torch.float32, tf32=True),
            "eager bf16ac": oracle("cuda", torch.float32, autocast=torch.bfloat16),
            "eager fp16ac": oracle("cuda", torch.float32, autocast=torch.float16),
            "pcadv fp32": ours("fp32"), "pcadv fp16": ours("fp16"), "fp16 fused": ours("fp16", fused=True),
            "pcadv bf16": ours("bf16")}
    names = list(arms)
    print("# relative L2 error of every parameter gradient of one adversarial step against float64; %d + %d "
          "clouds of %d points" % (B, B, N))
    print("%-22s %10s " % ("parameter", "|g|_2") + " ".join("%12s" % n for n in names))
    tot = {n: [0.0, 0.0] for n in names}
    for k, t in truth.items():
        nrm = t.norm().item()
        small = t.abs() < 1e-2 * t.abs().max()
        line = "%-22s %10.3e " % (k, nrm)
        errs, serr = [], []
        for n in names:
            e = arms[n][k] - t
            errs.append(e.norm().item() / max(nrm, 1e-300))
            # Adam-like sensitivity: error relative to the element's own magnitude, small elements only
            rel = (e.abs() / (t.abs() + 1e-12))[small]
            serr.append(rel.median().item() if rel.numel() else float("nan"))
            tot[n][0] += e.norm().item() ** 2
            tot[n][1] += nrm ** 2
        print(line + " ".join("%12.2e" % v for v in errs))
    print("%-22s %10s " % ("all", "") + " ".join("%12.2e" % ((tot[n][0] / tot[n][1]) ** 0.5) for n in names))


if __name__ == "__main__":
    main()

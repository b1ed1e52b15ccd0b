#!/usr/bin/env python3
"""Diagnostic (not collected by pytest): per-tensor branch-conditioned gradient errors of
PointNetDenseCls in the fp16 mode at several sizes, tensor-core vs CUDA-core engine, and the
per-tensor gradient difference between the one-pass and the two-pass fused step."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import test_gpu_branch_parity as T
    import test_gpu_graphed as TG
    from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step_fused
    from helpers import rel_err
    for engine in ("tc", "simt"):
        os.environ["PCADV_ENGINE"] = engine
        for B, N in ((3, 500), (3, 512), (8, 2500), (32, 2500), (32, 2560)):
            try:
                T._dense_parity("fp16", B, N, 50)
            except AssertionError as e:
                rep = e.args[0] if e.args and isinstance(e.args[0], dict) else None
                if rep is None:
                    print("densecls", engine, B, N, "failed:", str(e)[:300])
                    continue
                print("densecls %s %dx%d: total %.2e fwd %s" % (engine, B, N, rep["grad_total"], rep["fwd"]))
                print("   ", {k: "%.1e" % v for k, v in rep["grads"].items()})
            else:
                print("densecls %s %dx%d: passed" % (engine, B, N))
    os.environ["PCADV_ENGINE"] = "tc"
    for mode in ("fp32", "fp16"):
        B, N = 3, 320
        g, d = TG._models(N, mode, seed=7)
        g.to("cuda"); d.to("cuda")
        (pts, cls, seg), (pts2, cls2) = [tuple(t.cuda() for t in part) for part in TG._batches(B, N, 1)[0]]
        targs = argparse.Namespace(device="cuda", lambda_seg=1.0, lambda_adv=0.5)
        gan, ce = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()
        res = {}
        for tag, one_pass in (("two", False), ("one", True), ("two_again", False)):
            opt = torch.optim.SGD(g.parameters(), lr=0.0)
            optD = torch.optim.SGD(d.parameters(), lr=0.0)
            torch.manual_seed(5)
            l = adversarial_seg_step_fused(g, d, gan, ce, opt, optD, (pts, cls, seg), (pts2, cls2), targs,
                                           one_pass=one_pass)
            res[tag] = (torch.stack(l).cpu(), {("G." if i < 20 else "D.") + k: v.grad.clone() for i, (k, v) in
                                               enumerate(list(g.named_parameters()) + list(d.named_parameters()))})
        print(mode, "losses", res["two"][0].tolist(), res["one"][0].tolist())
        print(mode, "one vs two:", {k: "%.1e" % rel_err(res["one"][1][k], res["two"][1][k]) for k in res["one"][1]})
        print(mode, "two vs two:", {k: "%.1e" % rel_err(res["two_again"][1][k], res["two"][1][k]) for k in res["one"][1]})


if __name__ == "__main__":
    main()

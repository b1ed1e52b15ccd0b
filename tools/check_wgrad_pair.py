#!/usr/bin/env python3
"""Validate the CTA-pair weight-gradient kernel (PCADV_WGRAD_PAIR=1) against fp64 torch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from adversarial_learning_on_pointclouds_b200 import ops
from adversarial_learning_on_pointclouds_b200.ops import ENGINE_TC
dev = "cuda"
torch.manual_seed(0)
for rows, ks, rpg in [(4096, [64, 128, 128, 128, 512], 1024), (20000 - 16, [512], 0), (100000, [64, 128, 128, 128, 512], 0), (1 << 20, [64, 128, 128, 128, 512], 4096)]:
    n = 256
    dz = (torch.randn(rows, n, device=dev)).half()
    segs = [torch.randn(rows, k, device=dev).half() for k in ks]
    dw = torch.zeros(n, sum(ks), device=dev)
    dgb = torch.zeros(rows // rpg, n, device=dev) if rpg else None
    sc = torch.tensor([0.5], device=dev)
    ops.wgrad(dz, segs, dw=dw, dgroup_bias=dgb, rows_per_group=rpg, scale=sc, engine=ENGINE_TC)
    torch.cuda.synchronize()
    ref = 0.5 * dz.double().t() @ torch.cat(segs, 1).double()
    err = ((dw.double() - ref).norm() / ref.norm()).item()
    msg = "rows %d ks %s: dw rel err %.2e" % (rows, ks, err)
    if rpg:
        gref = dz.double().view(rows // rpg, rpg, n).sum(1)
        msg += "  dgroup rel err %.2e" % ((dgb.double() - gref).norm() / gref.norm()).item()
    print(msg)

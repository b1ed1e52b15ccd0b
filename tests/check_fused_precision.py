#!/usr/bin/env python3
"""Gradient error of the reference-shaped and the fused-head paths against the fp64-free CPU
oracle, per loss term (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from adversarial_learning_on_pointclouds_b200 import models as M, Precision
from adversarial_learning_on_pointclouds_b200.utils import init_net
from oracle import pointnet_oracle as PO, discriminator_oracle as DO, steps

dev = "cuda"
B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(0)
g = init_net(M.PointNetSeg(50), "cpu", "xavier")
d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
g.to(dev); d.to(dev)
pts, _, seg, cls = PO.synthetic_inputs(B, N, 1234)
P, C, S = pts.to(dev), cls.to(dev), seg.to(dev)

def rel(a, b):
    return ((a.detach().cpu().double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()

def oracle(kind):
    for p_ in list(gp.values()) + list(dp.values()):
        p_.grad = None
    o_pred, _ = PO.pointnet_seg_forward(gp, pts, cls)
    if kind == "ce":
        loss = F.cross_entropy(o_pred, seg)
    else:
        o_D = DO.pointwise_disc_forward(dp, F.log_softmax(o_pred, dim=1), N)
        loss = F.binary_cross_entropy_with_logits(o_D, torch.ones_like(o_D))
    loss.backward()
    return loss.item(), {k: v.grad.clone() for k, v in gp.items()}

for mode in ("fp32", "fp16"):
    g.precision = d.precision = Precision(mode)
    for kind in ("ce", "adv"):
        ol, og = oracle(kind)
        for p_ in d.parameters():
            p_.requires_grad = False
        # reference-shaped
        g.zero_grad()
        pred, _ = g(P, C)
        if kind == "ce":
            loss = F.cross_entropy(pred, S)
        else:
            Do = d(F.log_softmax(pred, dim=1)); loss = F.binary_cross_entropy_with_logits(Do, torch.ones_like(Do))
        loss.backward()
        e_ref = {k: rel(v.grad, og[k]) for k, v in g.named_parameters()}
        # fused
        g.zero_grad()
        if kind == "ce":
            loss2, _, _ = g.forward_ce(P, C, S)
        else:
            lp, _ = g.forward_logsoftmax(P, C); Do = d(lp)
            loss2 = F.binary_cross_entropy_with_logits(Do, torch.ones_like(Do))
        loss2.backward()
        e_fus = {k: rel(v.grad, og[k]) for k, v in g.named_parameters()}
        print("%s %s: loss oracle %.6f ref %.6f fused %.6f" % (mode, kind, ol, loss.item(), loss2.item()))
        worst = sorted(e_fus, key=lambda k: -e_fus[k])[:4]
        for k in worst:
            print("    %-14s ref-shaped %.2e   fused %.2e" % (k, e_ref[k], e_fus[k]))

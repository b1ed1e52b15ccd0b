#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (imported read-only from /root/reference) on CPU.

Run in the build container only (``python tests/golden/make_golden.py``); the
GPU box has no /root/reference and only reads the committed ``golden.pt``.
The reference holds no golden vectors of its own, so these
outputs of the reference itself are what pins the oracle (and through it the
CUDA path).  Shims, as in SURVEY.md D6: a ``matplotlib`` stub and
``np.object`` before importing ``utils.trainer``.

Every case stores: the recipe (seeds / sizes), a checksum of the weights (so a
test can prove it rebuilt the same parameters), scalar losses, summary
statistics and seeded probes of logits and of every parameter gradient.
"""
import argparse
import logging
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.switch_backend = lambda *a, **k: None
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    for name, val in (("object", object), ("bool", bool), ("int", int)):
        if not hasattr(np, name):
            setattr(np, name, val)


def inputs(B, N, seed):
    """SURVEY.md §8c ``inputs(B, N, seed)``."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(B, N, 3, generator=g) * 2 - 1
    y = torch.randint(0, 40, (B,), generator=g)
    seg = torch.randint(0, 50, (B, N), generator=g)
    shp = torch.randint(0, 16, (B,), generator=g)
    cls = F.one_hot(shp, 16).to(torch.float32).view(B, 1, 16)
    return pts, y, seg, cls


def probe_idx(numel, n=48, seed=99):
    g = torch.Generator().manual_seed(seed + numel)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def summarize(t):
    t = t.detach().to(torch.float32).contiguous().reshape(-1)
    idx = probe_idx(t.numel())
    return dict(norm=t.double().norm().item(), sum=t.double().sum().item(),
                abssum=t.double().abs().sum().item(), probe=t[idx].clone(), numel=t.numel())


def weight_checksum(module):
    return {k: (v.double().sum().item(), v.double().abs().sum().item())
            for k, v in module.state_dict().items()}


def grad_summary(module):
    return {k: summarize(p.grad) for k, p in module.named_parameters() if p.grad is not None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(HERE, "golden.pt"))
    args = ap.parse_args()
    sys.path.insert(0, REF)
    install_shims()
    torch.set_num_threads(8)
    from models.pointnet import (PointNetCls, PointNetSeg, PointNetSeg_regulization, STNkd, STN3d,
                                 PointNetDenseCls, PointNetfeat)
    from models import discriminator as RD
    from utils.model_utils import init_net
    import utils.trainer as RT
    from utils.image_pool import ImagePool

    G = {}

    # ---- KAT-1: PointNetCls eval, B=32 N=2500
    torch.manual_seed(0)
    m = PointNetCls(40, False); m.eval()
    pts, y, seg, cls = inputs(32, 2500, 1234)
    logits, glob, _ = m(pts)
    loss = F.cross_entropy(logits, y); loss.backward()
    with torch.no_grad():
        x = pts.transpose(1, 2)
        h = F.relu(m.feat.conv2(F.relu(m.feat.conv1(x))))
        h = m.feat.conv4(F.relu(m.feat.conv3(h)))
        amax = h.max(2)[1]
    G["kat1_cls"] = dict(recipe=dict(B=32, N=2500, seed=1234, wseed=0, k=40),
                         weights=weight_checksum(m), loss=loss.item(), logits=logits.detach().clone(),
                         glob=summarize(glob), argmax_sum=int(amax.sum().item()),
                         argmax=amax.to(torch.int32).clone(),
                         gradnorm=float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters()))),
                         grads=grad_summary(m))
    print("KAT-1", loss.item(), logits.abs().sum().item(), glob.sum().item(), amax.sum().item(),
          G["kat1_cls"]["gradnorm"])

    # ---- KAT-2: PointNetSeg xavier, B=16 N=2048
    torch.manual_seed(0)
    g = init_net(PointNetSeg(50), "cpu", "xavier")
    pts, y, seg, cls = inputs(16, 2048, 1234)
    pred, glob = g(pts, cls)
    loss = F.cross_entropy(pred, seg); loss.backward()
    G["kat2_seg"] = dict(recipe=dict(B=16, N=2048, seed=1234, wseed=0, k=50),
                         weights=weight_checksum(g), loss=loss.item(), pred=summarize(pred),
                         pred_stride=tuple(pred.stride()), glob=summarize(glob),
                         glob_zero=int((glob == 0).sum().item()),
                         gradnorm=float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in g.parameters()))),
                         grads=grad_summary(g))
    print("KAT-2", loss.item(), pred.abs().sum().item(), glob.sum().item(), (glob == 0).sum().item(),
          G["kat2_seg"]["gradnorm"], pred.stride())

    # ---- KAT-3: G-phase of the adversarial seg step, B=8+8 N=2048
    torch.manual_seed(0)
    g = init_net(PointNetSeg(50), "cpu", "xavier")
    d = init_net(RD.PointwiseDiscNet(2048, 50), "cpu", "xavier")
    for p in d.parameters():
        p.requires_grad = False
    pts, y, seg, cls = inputs(8, 2048, 1234)
    pts2, _, _, cls2 = inputs(8, 2048, 4321)
    pred, _ = g(pts, cls)
    l_seg = F.cross_entropy(pred, seg)
    pred2, _ = g(pts2, cls2)
    D_out = d(F.log_softmax(pred2, dim=1))
    l_adv = F.binary_cross_entropy_with_logits(D_out, torch.ones_like(D_out))
    (1.0 * l_seg + 0.001 * l_adv).backward()
    G["kat3_adv"] = dict(recipe=dict(B=8, N=2048, seed=1234, seed2=4321, wseed=0),
                         weights_g=weight_checksum(g), weights_d=weight_checksum(d),
                         l_seg=l_seg.item(), l_adv=l_adv.item(), dout=summarize(D_out),
                         gradnorm=float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in g.parameters()))),
                         grads=grad_summary(g), d_has_grads=any(p.grad is not None for p in d.parameters()))
    print("KAT-3", l_seg.item(), l_adv.item(), D_out.sum().item(), G["kat3_adv"]["gradnorm"])

    # ---- KAT-4: STNkd(64) + restated regulariser
    torch.manual_seed(0)
    s = STNkd(64)
    x = torch.rand(4, 64, 512, generator=torch.Generator().manual_seed(1234))
    t = s(x)
    eye = torch.eye(64)[None]
    reg = torch.mean(torch.norm(torch.bmm(t, t.transpose(2, 1)) - eye, dim=(1, 2)))
    reg.backward()
    G["kat4_stn"] = dict(recipe=dict(B=4, N=512, k=64, seed=1234, wseed=0), weights=weight_checksum(s),
                         trans=t.detach().clone(), reg=reg.item(), grads=grad_summary(s))
    print("KAT-4", t.sum().item(), reg.item())

    # ---- small full-tensor cases ------------------------------------------------
    def small_seg(name, regu):
        torch.manual_seed(3)
        net = init_net((PointNetSeg_regulization if regu else PointNetSeg)(50), "cpu", "xavier")
        with torch.no_grad():      # non-zero biases so bias paths are pinned too
            gb = torch.Generator().manual_seed(11)
            for n_, p in net.named_parameters():
                if n_.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=gb) * 0.05)
        pts, y, seg, cls = inputs(3, 200, 77)
        out = net(pts, cls)
        pred, glob = out[0], out[1]
        loss = F.cross_entropy(pred, seg) + 0.5 * glob.square().mean()
        if regu:
            tf = out[2]
            eye = torch.eye(128)[None]
            loss = loss + 1e-3 * torch.mean(torch.norm(torch.bmm(tf, tf.transpose(2, 1)) - eye, dim=(1, 2)))
        loss.backward()
        G[name] = dict(recipe=dict(B=3, N=200, seed=77, wseed=3, bseed=11), weights=weight_checksum(net),
                       loss=loss.item(), pred=pred.detach().clone(), glob=glob.detach().clone(),
                       grads=grad_summary(net))
        if regu:
            G[name]["trans_feat"] = summarize(out[2])
        print(name, loss.item())

    small_seg("small_seg", False)
    small_seg("small_seg_regu", True)

    # PointNetCls with feature transform + regulariser (cfg4 shape, small)
    torch.manual_seed(5)
    m = PointNetCls(40, True); m.eval()
    pts, y, seg, cls = inputs(4, 160, 55)
    logits, glob, tf = m(pts)
    eye = torch.eye(64)[None]
    reg = torch.mean(torch.norm(torch.bmm(tf, tf.transpose(2, 1)) - eye, dim=(1, 2)))
    loss = F.cross_entropy(logits, y) + 1e-3 * reg
    loss.backward()
    G["small_cls_ft"] = dict(recipe=dict(B=4, N=160, seed=55, wseed=5), weights=weight_checksum(m),
                             loss=loss.item(), reg=reg.item(), logits=logits.detach().clone(),
                             glob=glob.detach().clone(), trans=summarize(tf), grads=grad_summary(m))
    print("small_cls_ft", loss.item(), reg.item())

    # PointNetDenseCls with the two-line fix applied to the forward only
    torch.manual_seed(6)
    m = PointNetDenseCls(num_classes=50)

    def dense_fwd(self, x):
        batchsize, n_pts = x.size(0), x.size(2)
        x, trans_feat = self.feat(x)
        x = F.relu(self.conv1(x)); x = F.relu(self.conv2(x)); x = F.relu(self.conv3(x))
        x = self.conv4(x)
        x = x.transpose(2, 1).contiguous()
        x = F.log_softmax(x.view(-1, self.num_classes), dim=-1)
        return x.view(batchsize, n_pts, self.num_classes), trans_feat

    pts, y, seg, cls = inputs(3, 200, 66)
    out, _ = dense_fwd(m, pts.transpose(1, 2).contiguous())
    loss = F.nll_loss(out.reshape(-1, 50), seg.reshape(-1)); loss.backward()
    G["small_densecls"] = dict(recipe=dict(B=3, N=200, seed=66, wseed=6), weights=weight_checksum(m),
                               loss=loss.item(), out=out.detach().clone(), grads=grad_summary(m))
    print("small_densecls", loss.item())

    # discriminators, each on a B x 50 x N probability map
    def disc_case(name, ctor, run, wseed):
        torch.manual_seed(wseed)
        mods = ctor()
        mods = [init_net(mm, "cpu", "xavier") for mm in mods]
        with torch.no_grad():
            gb = torch.Generator().manual_seed(13)
            for mm in mods:
                for n_, p in mm.named_parameters():
                    if n_.endswith("bias"):
                        p.copy_(torch.randn(p.shape, generator=gb) * 0.05)
        gi = torch.Generator().manual_seed(21)
        x = torch.log_softmax(torch.randn(3, 50, 200, generator=gi), dim=1).requires_grad_(True)
        outs = run(mods, x)
        loss = sum((o * torch.linspace(0.5, 1.5, o.numel()).view_as(o)).mean() for o in outs)
        loss.backward()
        G[name] = dict(recipe=dict(B=3, N=200, wseed=wseed, bseed=13, xseed=21),
                       weights=[weight_checksum(mm) for mm in mods], loss=loss.item(),
                       outs=[o.detach().clone() for o in outs], dx=summarize(x.grad),
                       grads=[grad_summary(mm) for mm in mods])
        print(name, loss.item())

    disc_case("disc_pointwise", lambda: [RD.PointwiseDiscNet(200, 50)], lambda m, x: [m[0](x)], 31)
    disc_case("disc_conv", lambda: [RD.ConvDiscNet(50)], lambda m, x: [m[0](x.transpose(1, 2))], 32)
    disc_case("disc_stack", lambda: [RD.StackDiscNet(200, 50, 16)], lambda m, x: list(m[0](x)), 33)

    def dual_run(m, x):
        shared = m[0](x)
        return [m[1](shared), m[2](shared)]
    disc_case("disc_dual", lambda: [RD.BaseDiscNet(200, 50, 256), RD.ShapeDiscNet(256, 16),
                                    RD.PointDiscNet(256, 200)], dual_run, 34)

    torch.manual_seed(35)
    dd = init_net(RD.DeepConvDiscNet(40, 1), "cpu", "xavier")
    x = torch.log_softmax(torch.randn(6, 40, generator=torch.Generator().manual_seed(22)), 1).requires_grad_(True)
    o = dd(x); loss = (o * torch.linspace(0.5, 1.5, 6).view_as(o)).mean(); loss.backward()
    G["disc_deepconv"] = dict(recipe=dict(B=6, wseed=35, xseed=22), weights=[weight_checksum(dd)],
                              loss=loss.item(), outs=[o.detach().clone()], dx=summarize(x.grad),
                              grads=[grad_summary(dd)])
    print("disc_deepconv", loss.item())

    # ---- the unmodified trainer loop, one iteration (utils/trainer.py:849-1026) ----
    torch.manual_seed(0)
    g = init_net(PointNetSeg(50), "cpu", "xavier")
    d = init_net(RD.PointwiseDiscNet(256, 50), "cpu", "xavier")
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(d.parameters(), lr=1e-5, betas=(0.9, 0.999))
    pts, y, seg, cls = inputs(2, 256, 1234)
    pts2, _, _, cls2 = inputs(2, 256, 4321)
    # test set with all 16 categories
    tp, _, tseg, _ = inputs(16, 256, 999)
    tcls = F.one_hot(torch.arange(16), 16).float().view(16, 1, 16)
    testloader = [(tp[i:i + 4], tcls[i:i + 4], tseg[i:i + 4]) for i in range(0, 16, 4)]
    tmp = tempfile.mkdtemp()
    a = argparse.Namespace(device="cpu", total_iterations=1, iter_save_epoch=1, iter_test_epoch=1,
                           tensorboard=False, exp_dir=tmp, batch_size=2, input_pts=256,
                           lambda_seg=1.0, lambda_adv=1e-3)
    logger = logging.getLogger("golden"); logger.setLevel("ERROR")
    torch.manual_seed(4242)          # pins the two make_D_label(random=True) draws
    RT.run_training_seg([(pts, cls, seg)], [(pts2, cls2)], enumerate([(pts, cls, seg)]),
                        enumerate([(pts2, cls2)]), testloader, list(range(16)), g, d,
                        torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), opt, optD,
                        ImagePool(0), ImagePool(0), logger, logger, None, a)
    G["trainer_seg_step"] = dict(
        recipe=dict(B=2, N=256, seed=1234, seed2=4321, wseed=0, label_seed=4242, lr_g=1e-4, lr_d=1e-5),
        g_after={k: summarize(v) for k, v in g.state_dict().items()},
        d_after={k: summarize(v) for k, v in d.state_dict().items()},
        g_grads=grad_summary(g), d_grads=grad_summary(d))
    print("trainer_seg_step done")

    torch.save(G, args.out)
    print("wrote", args.out, os.path.getsize(args.out) / 1e6, "MB")


if __name__ == "__main__":
    main()

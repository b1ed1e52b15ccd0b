"""CPU: the oracle restatement against the golden outputs of the unmodified
reference (tests/golden/golden.pt, made by tests/golden/make_golden.py).  Also
proves that the package's module constructors + init_net rebuild exactly the
reference's parameters from a seed (weight checksums)."""
import torch
import torch.nn.functional as F

from adversarial_learning_on_pointclouds_b200 import models as M
from adversarial_learning_on_pointclouds_b200.utils import init_net
from oracle import pointnet_oracle as PO, discriminator_oracle as DO, steps
from helpers import (assert_summary_close, build_seg, check_weights, inputs, randomize_biases,
                     rel_err)

TOL = 2e-5      # fp32 CPU vs fp32 CPU, different op order at most


def _grads(params):
    return {k: v.grad for k, v in params.items() if v.grad is not None}


def test_kat1_pointnet_cls(golden):
    G = golden["kat1_cls"]
    torch.manual_seed(0)
    m = M.PointNetCls(40, False)
    check_weights(m, G["weights"])
    pts, y, _, _ = inputs(32, 2500, 1234)
    p = steps.leaf_params(m.state_dict())
    rec = {}
    logits, glob, tf = PO.pointnet_cls_forward(p, pts, record=rec)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    assert tf is None
    assert abs(loss.item() - G["loss"]) < 1e-6
    assert abs(loss.item() - 3.68183970) < 1e-6
    assert rel_err(logits, G["logits"]) < TOL
    assert_summary_close(glob, G["glob"], TOL, "glob")
    assert int(rec["argmax:feat.conv4"].sum()) == G["argmax_sum"] == 42048388
    for k, g in _grads(p).items():
        assert_summary_close(g, G["grads"][k], 1e-4, k)


def test_kat2_pointnet_seg(golden):
    G = golden["kat2_seg"]
    net = build_seg(0)
    check_weights(net, G["weights"])
    pts, _, seg, cls = inputs(16, 2048, 1234)
    p = steps.leaf_params(net.state_dict())
    pred, glob = PO.pointnet_seg_forward(p, pts, cls)
    loss = F.cross_entropy(pred, seg)
    loss.backward()
    assert tuple(pred.stride()) == G["pred_stride"] == (102400, 1, 50)
    assert abs(loss.item() - G["loss"]) < 1e-6 and abs(loss.item() - 3.91233063) < 1e-6
    assert_summary_close(pred, G["pred"], TOL, "pred")
    assert int((glob == 0).sum()) == G["glob_zero"] == 3974
    gn = torch.sqrt(sum((v.grad.double() ** 2).sum() for v in p.values())).item()
    assert abs(gn - G["gradnorm"]) < 1e-4 * G["gradnorm"]
    for k, g in _grads(p).items():
        assert_summary_close(g, G["grads"][k], 2e-3, k)     # ReLU / argmax ties move single entries


def test_kat3_adversarial_g_phase(golden):
    G = golden["kat3_adv"]
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(2048, 50), "cpu", "xavier")
    check_weights(g, G["weights_g"]); check_weights(d, G["weights_d"])
    pts, _, seg, cls = inputs(8, 2048, 1234)
    pts2, _, _, cls2 = inputs(8, 2048, 4321)
    gp = steps.leaf_params(g.state_dict())
    dp = {k: v.detach().clone() for k, v in d.state_dict().items()}
    pred, _ = PO.pointnet_seg_forward(gp, pts, cls)
    l_seg = F.cross_entropy(pred, seg)
    pred2, _ = PO.pointnet_seg_forward(gp, pts2, cls2)
    D_out = DO.pointwise_disc_forward(dp, F.log_softmax(pred2, dim=1), 2048)
    l_adv = F.binary_cross_entropy_with_logits(D_out, torch.ones_like(D_out))
    (l_seg + 0.001 * l_adv).backward()
    assert abs(l_seg.item() - G["l_seg"]) < 1e-6 and abs(l_adv.item() - G["l_adv"]) < 1e-6
    assert_summary_close(D_out, G["dout"], TOL, "D_out")
    gn = torch.sqrt(sum((v.grad.double() ** 2).sum() for v in gp.values())).item()
    assert abs(gn - G["gradnorm"]) < 1e-4 * G["gradnorm"]
    assert not G["d_has_grads"]


def test_kat4_stnkd_regulariser(golden):
    G = golden["kat4_stn"]
    torch.manual_seed(0)
    s = M.STNkd(64)
    check_weights(s, G["weights"])
    x = torch.rand(4, 64, 512, generator=torch.Generator().manual_seed(1234))
    p = steps.leaf_params(s.state_dict())
    t = PO.stn_forward(p, x, 64)
    reg = PO.feature_transform_regularizer(t)
    reg.backward()
    assert rel_err(t, G["trans"]) < TOL
    assert abs(reg.item() - G["reg"]) < 1e-5 and abs(reg.item() - 3.63242769) < 1e-5
    for k, g in _grads(p).items():
        assert_summary_close(g, G["grads"][k], 1e-4, k)


def test_small_seg_full_tensors(golden):
    for name, regu in (("small_seg", False), ("small_seg_regu", True)):
        G = golden[name]
        net = build_seg(3, 11, regu)
        check_weights(net, G["weights"])
        pts, _, seg, cls = inputs(3, 200, 77)
        p = steps.leaf_params(net.state_dict())
        out = PO.pointnet_seg_forward(p, pts, cls, regulization=regu)
        pred, glob = out[0], out[1]
        loss = F.cross_entropy(pred, seg) + 0.5 * glob.square().mean()
        if regu:
            loss = loss + 1e-3 * PO.feature_transform_regularizer(out[2])
        loss.backward()
        assert abs(loss.item() - G["loss"]) < 1e-6
        assert rel_err(pred, G["pred"]) < TOL and rel_err(glob, G["glob"]) < TOL
        for k, g in _grads(p).items():
            assert_summary_close(g, G["grads"][k], 1e-4, name + ":" + k)


def test_small_cls_feature_transform(golden):
    G = golden["small_cls_ft"]
    torch.manual_seed(5)
    m = M.PointNetCls(40, True)
    check_weights(m, G["weights"])
    pts, y, _, _ = inputs(4, 160, 55)
    p = steps.leaf_params(m.state_dict())
    logits, glob, tf = PO.pointnet_cls_forward(p, pts, feature_transform=True)
    reg = PO.feature_transform_regularizer(tf)
    loss = F.cross_entropy(logits, y) + 1e-3 * reg
    loss.backward()
    assert abs(loss.item() - G["loss"]) < 1e-6 and abs(reg.item() - G["reg"]) < 1e-5
    assert rel_err(logits, G["logits"]) < TOL and rel_err(glob, G["glob"]) < TOL
    for k, g in _grads(p).items():
        assert_summary_close(g, G["grads"][k], 1e-4, k)


def test_small_densecls_fixed(golden):
    G = golden["small_densecls"]
    torch.manual_seed(6)
    m = M.PointNetDenseCls(num_classes=50)
    check_weights(m, G["weights"])
    pts, _, seg, _ = inputs(3, 200, 66)
    p = steps.leaf_params(m.state_dict())
    out, _ = PO.pointnet_densecls_forward(p, pts.transpose(1, 2).contiguous(), 50)
    loss = F.nll_loss(out.reshape(-1, 50), seg.reshape(-1))
    loss.backward()
    assert abs(loss.item() - G["loss"]) < 1e-6
    assert rel_err(out, G["out"]) < TOL
    for k, g in _grads(p).items():
        assert_summary_close(g, G["grads"][k], 1e-4, k)


def _disc_case(G, mods, run, x):
    for mm, w in zip(mods, G["weights"]):
        check_weights(mm, w)
    ps = [steps.leaf_params(mm.state_dict()) for mm in mods]
    outs = run(ps, x)
    loss = sum((o * torch.linspace(0.5, 1.5, o.numel()).view_as(o)).mean() for o in outs)
    loss.backward()
    assert abs(loss.item() - G["loss"]) < 1e-5
    for o, ref in zip(outs, G["outs"]):
        assert rel_err(o, ref) < TOL
    assert_summary_close(x.grad, G["dx"], 1e-4, "dx")
    for p, gr in zip(ps, G["grads"]):
        for k, g in _grads(p).items():
            assert_summary_close(g, gr[k], 1e-4, k)


def _disc_mods(ctor, wseed):
    torch.manual_seed(wseed)
    mods = [init_net(mm, "cpu", "xavier") for mm in ctor()]
    randomize_biases(mods, 13)
    return mods


def _disc_x():
    gi = torch.Generator().manual_seed(21)
    return torch.log_softmax(torch.randn(3, 50, 200, generator=gi), dim=1).requires_grad_(True)


def test_discriminators(golden):
    _disc_case(golden["disc_pointwise"], _disc_mods(lambda: [M.PointwiseDiscNet(200, 50)], 31),
               lambda p, x: [DO.pointwise_disc_forward(p[0], x, 200)], _disc_x())
    _disc_case(golden["disc_conv"], _disc_mods(lambda: [M.ConvDiscNet(50)], 32),
               lambda p, x: [DO.conv_disc_forward(p[0], x.transpose(1, 2))], _disc_x())
    _disc_case(golden["disc_stack"], _disc_mods(lambda: [M.StackDiscNet(200, 50, 16)], 33),
               lambda p, x: list(DO.stack_disc_forward(p[0], x)), _disc_x())

    def dual(p, x):
        shared = DO.base_disc_forward(p[0], x)
        return [DO.shape_disc_forward(p[1], shared), DO.point_disc_forward(p[2], shared, 200)]
    _disc_case(golden["disc_dual"],
               _disc_mods(lambda: [M.BaseDiscNet(200, 50, 256), M.ShapeDiscNet(256, 16),
                                   M.PointDiscNet(256, 200)], 34), dual, _disc_x())
    G = golden["disc_deepconv"]
    torch.manual_seed(35)
    dd = init_net(M.DeepConvDiscNet(40, 1), "cpu", "xavier")
    x = torch.log_softmax(torch.randn(6, 40, generator=torch.Generator().manual_seed(22)), 1).requires_grad_(True)
    _disc_case(G, [dd], lambda p, x_: [DO.deepconv_disc_forward(p[0], x_)], x)


def test_trainer_seg_step(golden):
    """oracle.steps.adversarial_seg_step + Adam against one iteration of the
    reference's unmodified run_training_seg (utils/trainer.py:873-966)."""
    G = golden["trainer_seg_step"]
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(256, 50), "cpu", "xavier")
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    opt = torch.optim.Adam(list(gp.values()), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(list(dp.values()), lr=1e-5, betas=(0.9, 0.999))
    pts, _, seg, cls = inputs(2, 256, 1234)
    pts2, _, _, cls2 = inputs(2, 256, 4321)
    torch.manual_seed(4242)
    steps.adversarial_seg_step(gp, dp, (pts, cls, seg), (pts2, cls2))
    for k, v in gp.items():
        assert_summary_close(v.grad, G["g_grads"][k], 1e-4, "g:" + k)
    for k, v in dp.items():
        assert_summary_close(v.grad, G["d_grads"][k], 1e-4, "d:" + k)
    opt.step(); optD.step()
    for k, v in gp.items():
        assert_summary_close(v, G["g_after"][k], 1e-6, "g_after:" + k)
    for k, v in dp.items():
        assert_summary_close(v, G["d_after"][k], 1e-6, "d_after:" + k)


def test_bench_synthetic_inputs_match_the_oracle_generator():
    """bench.py carries its own copy of the SURVEY 8c input recipe (its CUDA arm must not import
    oracle/); it has to draw exactly what the oracle's generator draws."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from oracle.pointnet_oracle import synthetic_inputs
    for a, b in zip(bench.synthetic_inputs(3, 17, 1234), synthetic_inputs(3, 17, 1234)):
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------
# the remaining live loop bodies (tests/golden/golden_steps.pt, made by make_golden_steps.py from the
# reference's unmodified run_training_seg_dual / run_training_semi / run_training_seg_semi)
import os

import pytest


@pytest.fixture(scope="module")
def golden_steps():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_steps.pt"), weights_only=False)


def _check_after(params, expected, tol, tag):
    for k, v in params.items():
        assert_summary_close(v, expected[k], tol, tag + ":" + k)


def dual_setup(N):
    """The models of load_models("seg") / load_models("disc_dual") (utils/model_utils.py:81, :116-131)
    from seed 0, in the golden script's construction order."""
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    shared = init_net(M.BaseDiscNet(N, 50, 256), "cpu", "xavier")
    shape = init_net(M.ShapeDiscNet(256, 16), "cpu", "xavier")
    point = init_net(M.PointDiscNet(256, N), "cpu", "xavier")
    return g, shared, shape, point


def test_trainer_seg_dual_steps(golden_steps):
    """oracle.steps.adversarial_seg_dual_step against two iterations of the reference's unmodified
    run_training_seg_dual: three discriminators, optimizer_D_point never zeroed (utils/trainer.py:2171-2172),
    sharedDisc stepped by both discriminator optimizers."""
    G = golden_steps["dual"]
    R = G["recipe"]
    g, shared, shape, point = dual_setup(R["N"])
    gp, sp, hp, pp = (steps.leaf_params(m.state_dict()) for m in (g, shared, shape, point))
    opt = torch.optim.Adam(list(gp.values()), lr=R["lr_g"], betas=(0.9, 0.999))
    opt_shape = torch.optim.SGD(list(hp.values()) + list(sp.values()), lr=R["lr_d"])
    opt_point = torch.optim.SGD(list(pp.values()) + list(sp.values()), lr=R["lr_d"])
    torch.manual_seed(R["label_seed"])
    for it in range(R["iters"]):
        pts, _, seg, cls = inputs(R["B"], R["N"], R["seed"] + it)
        pts2, _, _, cls2 = inputs(R["B"], R["N"], R["seed2"] + it)
        steps.adversarial_seg_dual_step(gp, sp, hp, pp, (pts, cls, seg), (pts2, cls2), opt, opt_shape, opt_point,
                                        lambda_adv=R["lambda_adv"])
    _check_after(gp, G["g"], 1e-6, "g")
    _check_after(sp, G["shared"], 1e-6, "shared")
    _check_after(hp, G["shape"], 1e-6, "shape")
    _check_after(pp, G["point"], 1e-6, "point")


def test_trainer_cls_semi_steps(golden_steps):
    """oracle.steps.adversarial_cls_semi_step against three iterations of the reference's unmodified
    run_training_semi (the third one with the semi-supervised term, utils/trainer.py:727-739)."""
    G = golden_steps["cls_semi"]
    R = G["recipe"]
    torch.manual_seed(0)
    g = M.PointNetCls(40, False)
    d = init_net(M.DeepConvDiscNet(40, 1), "cpu", "xavier")
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    opt = torch.optim.Adam(list(gp.values()), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(list(dp.values()), lr=1e-5, betas=(0.9, 0.999))
    torch.manual_seed(R["label_seed"])
    saw_semi = False
    for it in range(R["iters"]):
        pts, y, _, _ = inputs(R["B"], R["N"], R["seed"] + it)
        pts2 = inputs(R["B"], R["N"], R["seed2"] + it)[0]
        r = steps.adversarial_cls_semi_step(gp, dp, (pts, y), (pts2,), opt, optD, it, R["semi_start"], R["semi_TH"])
        saw_semi |= r["l_semi"] is not None
    assert saw_semi
    _check_after(gp, G["g"], 1e-6, "g")
    _check_after(dp, G["d"], 1e-6, "d")


def test_trainer_seg_semi_steps(golden_steps):
    """oracle.steps.adversarial_seg_semi_step against three iterations of the reference's unmodified
    run_training_seg_semi (utils/trainer.py:1903-2121; its semi branch runs on the CPU only: it
    indexes a CPU tensor with a device mask, :2002)."""
    G = golden_steps["seg_semi"]
    R = G["recipe"]
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(R["N"], 50), "cpu", "xavier")
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    opt = torch.optim.Adam(list(gp.values()), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(list(dp.values()), lr=1e-5, betas=(0.9, 0.999))
    torch.manual_seed(R["label_seed"])
    saw_semi = False
    for it in range(R["iters"]):
        pts, _, seg, cls = inputs(R["B"], R["N"], R["seed"] + it)
        pts2, _, _, cls2 = inputs(R["B"], R["N"], R["seed2"] + it)
        r = steps.adversarial_seg_semi_step(gp, dp, (pts, cls, seg), (pts2, cls2), opt, optD, it, R["semi_start"],
                                            R["semi_TH"])
        saw_semi |= r["l_semi"] is not None
    assert saw_semi
    _check_after(gp, G["g"], 1e-6, "g")
    _check_after(dp, G["d"], 1e-6, "d")

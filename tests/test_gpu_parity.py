"""GPU: the CUDA path (through the C ABI) against the oracle, on the same seeded
inputs, and against the committed golden fixtures.

Tolerances (BASELINE.json north_star): 1e-5 relative in the fp32-accumulate
verification mode with max-pool argmax indices bit-exact (outside exact
near-ties, which are enumerated and bounded); 1e-3 relative on tensor cores
(fp16 operands, fp32 accumulate).  "relative" = ||a - b||_2 / ||b||_2 per tensor.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import adversarial_learning_on_pointclouds_b200 as pkg
from adversarial_learning_on_pointclouds_b200 import models as M, ops
from adversarial_learning_on_pointclouds_b200.ops import (ACT_LEAKY, ACT_NONE, ACT_RELU, ENGINE_SIMT,
                                                          Precision)
from adversarial_learning_on_pointclouds_b200.utils import init_net, make_D_label
from oracle import pointnet_oracle as PO, discriminator_oracle as DO, steps
from helpers import (assert_summary_close, build_seg, check_weights, inputs, randomize_biases,
                     rel_err)
import parity

DEV = "cuda"
TOL = {"fp32": 1e-5, "fp16": 1e-3}
MODES = ["fp32", "fp16"]


def _rand(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).to(DEV)


# ----------------------------------------------------------------------------- ops
@pytest.mark.parametrize("rows,ks,n,rpg", [(300, [3], 64, 100), (5000, [64, 128], 50, 2500),
                                           (257, [50], 64, 257), (1024, [128], 256, 256),
                                           (7, [1024], 9, 7)])
def test_linear_simt_matches_torch(rows, ks, n, rpg):
    segs = [_rand((rows, k), 10 + i) for i, k in enumerate(ks)]
    w = _rand((n, sum(ks)), 3, 0.1)
    bias = _rand((n,), 4)
    gb = _rand((rows // rpg, n), 5)
    add = _rand((rows, n), 6)
    mask = _rand((rows, n), 7)
    out, ckey, rkey = ops.linear(segs, w, bias=bias, group_bias=gb, rows_per_group=rpg, addend=add,
                                 act=ACT_LEAKY, slope=0.2, mask=mask, mask_act=ACT_RELU,
                                 colmax=True, rowmax=True)
    x = torch.cat(segs, 1)
    pre = x.double() @ w.double().t() + bias.double() + gb.double().repeat_interleave(rpg, 0) + add.double()
    ref = F.leaky_relu(pre, 0.2) * (mask > 0)
    assert rel_err(out, ref) < 1e-5
    cval, cidx = ops.max_finalize(ckey, ACT_NONE)
    rv, ri = pre.view(rows // rpg, rpg, n).max(1)
    assert rel_err(cval, rv) < 1e-5
    agree = (cidx.long() == ri).double().mean().item()
    assert agree > 0.995
    rval, ridx = ops.max_finalize(rkey, ACT_RELU)
    rv2, ri2 = F.relu(pre).max(1)
    assert rel_err(rval, rv2) < 1e-5
    assert ((ridx.long() == ri2) | (rv2 == 0)).double().mean().item() > 0.995


def test_linear_half_storage_and_scale():
    rows, k, n = 1000, 64, 128
    x = _rand((rows, k), 1).half()
    w = _rand((n, k), 2, 0.1)
    sc = torch.tensor([0.25], device=DEV)
    out, _, _ = ops.linear([x], w, act=ACT_RELU, out_dtype=torch.float16, out_scale=sc)
    ref = F.relu(x.double() @ w.double().t()) * 0.25
    assert out.dtype == torch.float16 and rel_err(out, ref) < 1e-3


@pytest.mark.parametrize("rows,ks,n,rpg", [(5000, [64, 128], 50, 2500), (300, [3], 64, 100),
                                           (40000, [128], 256, 4000)])
def test_wgrad_simt_matches_torch(rows, ks, n, rpg):
    dz = _rand((rows, n), 1)
    segs = [_rand((rows, k), 10 + i) for i, k in enumerate(ks)]
    sc = torch.tensor([0.5], device=DEV)
    dw = torch.zeros((n, sum(ks)), device=DEV)
    db = torch.zeros((n,), device=DEV)
    dgb = torch.zeros((rows // rpg, n), device=DEV)
    ops.wgrad(dz, segs, dw=dw, dbias=db, dgroup_bias=dgb, rows_per_group=rpg, scale=sc)
    x = torch.cat(segs, 1).double()
    assert rel_err(dw, 0.5 * dz.double().t() @ x) < 1e-5
    assert rel_err(db, 0.5 * dz.double().sum(0)) < 1e-5
    assert rel_err(dgb, dz.double().view(rows // rpg, rpg, n).sum(1)) < 1e-5


def test_maxpool_bwd_matches_autograd():
    B, N, k, n = 3, 500, 128, 256
    x = _rand((B * N, k), 1)
    w = _rand((n, k), 2, 0.1)
    b = _rand((n,), 3)
    _, key, _ = ops.linear([x], w, bias=b, want_out=False, colmax=True, rows_per_group=N)
    g, idx = ops.max_finalize(key, ACT_RELU)
    dg = _rand((B, n), 4)
    dw = torch.zeros((n, k), device=DEV); db = torch.zeros((n,), device=DEV)
    dx = torch.zeros((B * N, k), device=DEV)
    ops.maxpool_bwd(dg, g, idx, x, w, N, act=ACT_RELU, dw=dw, dbias=db, dx_acc=dx)
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    y = F.relu(xr @ wr.t() + br).view(B, N, n)
    gr = torch.gather(y, 1, idx.long().unsqueeze(1)).squeeze(1)
    (gr * dg.double()).sum().backward()
    assert rel_err(g, gr) < 1e-5
    assert rel_err(dw, wr.grad) < 1e-5 and rel_err(db, br.grad) < 1e-5 and rel_err(dx, xr.grad) < 1e-5
    # in-place form: dz += relu'(x) * dx, touching argmax rows only
    for dt, tol in ((torch.float32, 1e-5), (torch.float16, 1e-3)):
        dz0 = _rand((B * N, k), 9).to(dt)
        dz = dz0.clone()
        ops.maxpool_bwd(dg, g, idx, x.to(dt), w.to(dt), N, act=ACT_RELU, dz_inout=dz, prev_act=ACT_RELU)
        ref = dz0.double() + (x.to(dt) > 0) * xr.grad
        assert rel_err(dz, ref) < tol
        touched = (xr.grad.abs().sum(1) > 0)
        assert torch.equal(dz[~touched], dz0[~touched])


def test_amax_scale_and_convert():
    x = _rand((1000, 50), 1, 3e-7)
    s2 = ops.amax_scale(x, target=256.0)
    amax = x.abs().max().item()
    S = s2[0].item()
    assert S == 2.0 ** torch.floor(torch.log2(torch.tensor(256.0 / amax))).item()
    assert abs(s2[1].item() * S - 1.0) < 1e-7
    y = ops.convert(x, torch.float16, cols_pad=64, scale=s2[0:1])
    assert y.shape == (1000, 64) and (y[:, 50:] == 0).all()
    assert rel_err(y[:, :50], x.double() * S) < 1e-3
    assert ops.amax_scale(torch.zeros((4, 4), device=DEV))[0].item() == 1.0


# --------------------------------------------------------------------- PointNetSeg
def _seg_on_gpu(wseed, bseed, mode):
    net = build_seg(wseed, bseed).to(DEV)
    net.precision = Precision(mode)
    return net


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("B,N", [(3, 200), (2, 1024), (1, 1)])
def test_seg_matches_oracle(mode, B, N):
    net = _seg_on_gpu(3, 11, mode)
    pts, _, seg, cls = inputs(B, N, 77)
    rep = parity.seg_parity(net, pts.to(DEV), cls.to(DEV), seg.to(DEV), TOL[mode], mode=mode)
    print(mode, B, N, rep["pred"], rep["grad_total"], rep["n_branch_diff"],
          "worst flipped margins: act %.2f u, argmax %.2f u" % (rep["worst_act_margin_u"],
                                                               rep["worst_argmax_margin_u"]))
    if mode == "fp32":
        assert rep["n_branch_diff"] <= 2          # exact-rounding ties only


def test_seg_small_matches_reference_golden(golden):
    """Direct comparison with the stored outputs of the reference itself."""
    G = golden["small_seg"]
    net = _seg_on_gpu(3, 11, "fp32")
    check_weights(net, G["weights"])
    pts, _, seg, cls = inputs(3, 200, 77)
    pred, glob = net(pts.to(DEV), cls.to(DEV))
    loss = F.cross_entropy(pred, seg.to(DEV)) + 0.5 * glob.square().mean()
    loss.backward()
    assert abs(loss.item() - G["loss"]) < 1e-5 * abs(G["loss"])
    assert rel_err(pred, G["pred"]) < 1e-5 and rel_err(glob, G["glob"]) < 1e-5
    for k, v in net.named_parameters():
        assert_summary_close(v.grad, G["grads"][k], 1e-4, k)


def test_seg_kat2_reference_golden(golden):
    """KAT-2 of SURVEY.md 8c: B=16, N=2048, the reference's known answers."""
    G = golden["kat2_seg"]
    net = _seg_on_gpu(0, None, "fp32")
    check_weights(net, G["weights"])
    pts, _, seg, cls = inputs(16, 2048, 1234)
    pred, glob = net(pts.to(DEV), cls.to(DEV))
    loss = F.cross_entropy(pred, seg.to(DEV))
    loss.backward()
    assert tuple(pred.stride()) == (102400, 1, 50)
    assert abs(loss.item() - 3.91233063) < 2e-5
    assert_summary_close(pred, G["pred"], 1e-5, "pred")
    assert_summary_close(glob, G["glob"], 1e-5, "glob")
    assert abs(int((glob == 0).sum()) - 3974) <= 2
    gn = torch.sqrt(sum((v.grad.double() ** 2).sum() for v in net.parameters())).item()
    assert abs(gn - G["gradnorm"]) < 1e-3 * G["gradnorm"]


def test_seg_no_grad_eval_and_frozen_params():
    net = _seg_on_gpu(3, 11, "fp32")
    pts, _, seg, cls = inputs(2, 128, 5)
    with torch.set_grad_enabled(False):                    # utils/trainer.py:96
        pred, _ = net.eval()(pts.to(DEV), cls.to(DEV))
    assert not pred.requires_grad
    net.train()
    for p in net.parameters():
        p.requires_grad = False
    net.fc4.weight.requires_grad = True
    pred, _ = net(pts.to(DEV), cls.to(DEV))
    F.cross_entropy(pred, seg.to(DEV)).backward()
    assert net.fc4.weight.grad is not None and net.conv1.weight.grad is None


def test_cpu_tensors_fail_loudly():
    net = build_seg(3)
    pts, _, _, cls = inputs(1, 16, 1)
    with pytest.raises(RuntimeError):
        net(pts, cls)


# ------------------------------------------------------------ PointNetCls / DenseCls
def _grad_check(named_params, oracle_params, tol):
    errs = {k: rel_err(v.grad, oracle_params[k].grad) for k, v in named_params}
    assert max(errs.values()) <= tol, errs
    return errs


@pytest.mark.parametrize("mode", MODES)
def test_cls_kat1(golden, mode):
    G = golden["kat1_cls"]
    torch.manual_seed(0)
    m = M.PointNetCls(40, False).to(DEV).eval()
    m.precision = m.feat.precision = Precision(mode)
    check_weights(m, G["weights"])
    pts, y, _, _ = inputs(32, 2500, 1234)
    logits, glob, tf = m(pts.to(DEV))
    loss = F.cross_entropy(logits, y.to(DEV))
    loss.backward()
    tol = TOL[mode]
    assert tf is None and tuple(glob.shape) == (32, 1024, 1)
    assert abs(loss.item() - 3.68183970) < 10 * tol
    assert rel_err(logits, G["logits"]) < 4 * tol
    assert_summary_close(glob, G["glob"], tol, "glob")
    if mode == "fp32":
        gn = torch.sqrt(sum((v.grad.double() ** 2).sum() for v in m.parameters())).item()
        assert abs(gn - G["gradnorm"]) < 1e-3 * G["gradnorm"]


@pytest.mark.parametrize("mode", MODES)
def test_cls_feature_transform_golden(golden, mode):
    G = golden["small_cls_ft"]
    torch.manual_seed(5)
    m = M.PointNetCls(40, True).to(DEV).eval()
    for mod in m.modules():
        mod.precision = Precision(mode)
    check_weights(m, G["weights"])
    pts, y, _, _ = inputs(4, 160, 55)
    logits, glob, tf = m(pts.to(DEV))
    reg = M.feature_transform_regularizer(tf)
    loss = F.cross_entropy(logits, y.to(DEV)) + 1e-3 * reg
    loss.backward()
    tol = TOL[mode]
    assert abs(reg.item() - G["reg"]) < 10 * tol * G["reg"]
    assert rel_err(logits, G["logits"]) < 4 * tol and rel_err(glob, G["glob"]) < 4 * tol
    if mode == "fp32":
        for k, v in m.named_parameters():
            assert_summary_close(v.grad, G["grads"][k], 2e-4, k)


@pytest.mark.parametrize("mode", MODES)
def test_densecls_golden(golden, mode):
    G = golden["small_densecls"]
    torch.manual_seed(6)
    m = M.PointNetDenseCls(num_classes=50).to(DEV)
    for mod in m.modules():
        mod.precision = Precision(mode)
    check_weights(m, G["weights"])
    pts, _, seg, _ = inputs(3, 200, 66)
    out, tf = m(pts.transpose(1, 2).contiguous().to(DEV))
    loss = F.nll_loss(out.reshape(-1, 50), seg.reshape(-1).to(DEV))
    loss.backward()
    tol = TOL[mode]
    assert tuple(out.shape) == (3, 200, 50)
    assert rel_err(out, G["out"]) < 4 * tol
    if mode == "fp32":
        for k, v in m.named_parameters():
            assert_summary_close(v.grad, G["grads"][k], 2e-4, k)


# ------------------------------------------------------------------ discriminators
def _disc_mods(ctor, wseed, mode):
    torch.manual_seed(wseed)
    mods = [init_net(mm, "cpu", "xavier") for mm in ctor()]
    randomize_biases(mods, 13)
    for mm in mods:
        mm.to(DEV)
        mm.precision = Precision(mode)
    return mods


def _disc_check(G, mods, run, x, mode):
    for mm, w in zip(mods, G["weights"]):
        check_weights(mm, w)
    outs = run(mods, x)
    loss = sum((o * torch.linspace(0.5, 1.5, o.numel(), device=DEV).view_as(o)).mean() for o in outs)
    loss.backward()
    tol = TOL[mode]
    for o, ref in zip(outs, G["outs"]):
        assert tuple(o.shape) == tuple(ref.shape)
        assert rel_err(o, ref) < 4 * tol
    if mode == "fp32":
        assert_summary_close(x.grad, G["dx"], 2e-4, "dx")
        for mm, gr in zip(mods, G["grads"]):
            for k, v in mm.named_parameters():
                if k in gr:
                    assert_summary_close(v.grad, gr[k], 2e-4, k)
                else:
                    assert v.grad is None, k          # BaseDiscNet.conv4 (never applied)


@pytest.mark.parametrize("mode", MODES)
def test_discriminators_golden(golden, mode):
    def x50():
        gi = torch.Generator().manual_seed(21)
        return torch.log_softmax(torch.randn(3, 50, 200, generator=gi), dim=1).to(DEV).requires_grad_(True)
    _disc_check(golden["disc_pointwise"], _disc_mods(lambda: [M.PointwiseDiscNet(200, 50)], 31, mode),
                lambda m, x: [m[0](x)], x50(), mode)
    _disc_check(golden["disc_conv"], _disc_mods(lambda: [M.ConvDiscNet(50)], 32, mode),
                lambda m, x: [m[0](x.transpose(1, 2))], x50(), mode)
    _disc_check(golden["disc_stack"], _disc_mods(lambda: [M.StackDiscNet(200, 50, 16)], 33, mode),
                lambda m, x: list(m[0](x)), x50(), mode)

    def dual(m, x):
        shared = m[0](x)
        return [m[1](shared), m[2](shared)]
    _disc_check(golden["disc_dual"],
                _disc_mods(lambda: [M.BaseDiscNet(200, 50, 256), M.ShapeDiscNet(256, 16),
                                    M.PointDiscNet(256, 200)], 34, mode), dual, x50(), mode)
    torch.manual_seed(35)
    dd = init_net(M.DeepConvDiscNet(40, 1), "cpu", "xavier").to(DEV)
    dd.precision = Precision(mode)
    x = torch.log_softmax(torch.randn(6, 40, generator=torch.Generator().manual_seed(22)), 1)
    _disc_check(golden["disc_deepconv"], [dd], lambda m, x_: [m[0](x_)],
                x.to(DEV).requires_grad_(True), mode)


# ------------------------------------------------------------- the adversarial step
def _adv_step(g, d, opt, optD, batch_gt, batch_nogt, lambda_seg=1.0, lambda_adv=1e-3):
    """The loop body of utils/trainer.py:873-966, as the unmodified trainer runs it
    on the modules (history pools of size 0)."""
    gan_loss, seg_loss = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()
    g.train(); d.train()
    opt.zero_grad(); optD.zero_grad()
    for p in d.parameters():
        p.requires_grad = False
    pts, cls, seg = (t.to(DEV) for t in batch_gt)
    pred, _ = g(pts, cls)
    l_seg = seg_loss(pred, seg)
    pred_gt_softmax = F.softmax(pred, dim=1)
    pts2, cls2 = (t.to(DEV) for t in batch_nogt)
    pred2, _ = g(pts2, cls2)
    pred2_ls = F.log_softmax(pred2, dim=1)
    D_out = d(pred2_ls)
    l_adv = gan_loss(D_out, make_D_label(D_out, 1, DEV, random=False))
    (lambda_seg * l_seg + lambda_adv * l_adv).backward()
    for p in d.parameters():
        p.requires_grad = True
    D_out = d(pred_gt_softmax.detach())
    (gan_loss(D_out, make_D_label(D_out, 1, DEV, random=True)) * 0.5).backward()
    D_out = d(pred2_ls.detach())
    (gan_loss(D_out, make_D_label(D_out, 0, DEV, random=True)) * 0.5).backward()
    return l_seg.item(), l_adv.item()


def test_adversarial_step_matches_reference_trainer(golden):
    """One iteration of the reference's unmodified run_training_seg (golden) vs the
    same loop body driving the CUDA modules: gradients and the Adam-updated
    parameters of G and D."""
    G = golden["trainer_seg_step"]
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier").to(DEV)
    d = init_net(M.PointwiseDiscNet(256, 50), "cpu", "xavier").to(DEV)
    g.precision = d.precision = Precision("fp32")
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(d.parameters(), lr=1e-5, betas=(0.9, 0.999))
    pts, _, seg, cls = inputs(2, 256, 1234)
    pts2, _, _, cls2 = inputs(2, 256, 4321)
    torch.manual_seed(4242)
    _adv_step(g, d, opt, optD, (pts, cls, seg), (pts2, cls2))
    for k, v in g.named_parameters():
        assert_summary_close(v.grad, G["g_grads"][k], 2e-4, "g:" + k)
    for k, v in d.named_parameters():
        assert_summary_close(v.grad, G["d_grads"][k], 2e-4, "d:" + k)
    opt.step(); optD.step()
    for k, v in g.state_dict().items():
        assert_summary_close(v, G["g_after"][k], 1e-5, "g_after:" + k)
    for k, v in d.state_dict().items():
        assert_summary_close(v, G["d_after"][k], 1e-5, "d_after:" + k)


@pytest.mark.parametrize("mode", MODES)
def test_adversarial_step_vs_oracle_step(mode):
    """G-phase + D-phase gradients against oracle.steps on a second size."""
    torch.manual_seed(1)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(384, 50), "cpu", "xavier")
    randomize_biases([g, d], 3)
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    g.to(DEV); d.to(DEV)
    g.precision = d.precision = Precision(mode)
    pts, _, seg, cls = inputs(3, 384, 8)
    pts2, _, _, cls2 = inputs(3, 384, 9)
    opt = torch.optim.SGD(g.parameters(), lr=0.0)
    optD = torch.optim.SGD(d.parameters(), lr=0.0)
    torch.manual_seed(77)
    l_seg, l_adv = _adv_step(g, d, opt, optD, (pts, cls, seg), (pts2, cls2), lambda_adv=0.5)
    torch.manual_seed(77)
    ref = steps.adversarial_seg_step(gp, dp, (pts, cls, seg), (pts2, cls2), lambda_adv=0.5)
    tol = TOL[mode]
    assert abs(l_seg - ref["l_seg"]) < 10 * tol and abs(l_adv - ref["l_adv"]) < 10 * tol
    # unconditioned comparison: ReLU / argmax ties may move single entries, so the bound is
    # looser than the branch-conditioned one in test_seg_matches_oracle
    loose = max(50 * tol, 5e-3)
    for k, v in d.named_parameters():
        assert rel_err(v.grad, dp[k].grad) < loose, k
    for k, v in g.named_parameters():
        assert rel_err(v.grad, gp[k].grad) < loose, k


# ------------------------------------------------------------- classification loop bodies (a15)
@pytest.mark.parametrize("ft", [False, True])
@pytest.mark.parametrize("mode", MODES)
def test_pointnet_cls_step_vs_oracle(mode, ft):
    """trainer.pointnet_cls_step (utils/trainer.py:236-269) against oracle.steps.pointnet_cls_step:
    CE + lambda_regu * regulariser, same losses and gradients.  Dropout is switched off on both
    sides (its random stream cannot be shared between the CPU oracle and the device)."""
    from adversarial_learning_on_pointclouds_b200.trainer import pointnet_cls_step
    import argparse
    torch.manual_seed(3)
    m = M.PointNetCls(40, ft)
    randomize_biases([m], 6)
    gp = steps.leaf_params(m.state_dict())
    m.to(DEV)
    m.dropout.p = 0.0
    for mod in m.modules():
        mod.precision = Precision(mode)
    pts, y, _, _ = inputs(4, 300, 21)
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    targs = argparse.Namespace(device=DEV, lambda_cls=1.0, lambda_regu=1e-3)
    l_cls, l_regu = pointnet_cls_step(m, torch.nn.CrossEntropyLoss(), opt, (pts.to(DEV), y.to(DEV)), targs)
    ref = steps.pointnet_cls_step(gp, (pts, y), feature_transform=ft, lambda_regu=1e-3)
    tol = TOL[mode]
    assert abs(l_cls.item() - ref["l_cls"]) < 10 * tol
    assert (l_regu is None) == (not ft)
    if ft:
        assert abs(l_regu.item() - ref["l_regu"]) < 10 * tol * ref["l_regu"]
    # un-conditioned: with 4 x 300 points behind a 1024-channel max-pool one arg-max that flips in
    # the 16-bit forward moves the first layers' gradients by several per cent (DESIGN.md section 5;
    # the branch-conditioned comparison is test_cls_kat1 / test_seg_matches_oracle)
    loose = 5e-3 if mode == "fp32" else 0.15
    for k, v in m.named_parameters():
        assert rel_err(v.grad, gp[k].grad) < loose, k


@pytest.mark.parametrize("mode", MODES)
def test_adversarial_cls_step_vs_oracle(mode):
    """trainer.adversarial_cls_step (utils/trainer.py:426-559: PointNetCls against DeepConvDiscNet on
    log_softmax maps) against oracle.steps.adversarial_cls_step."""
    from adversarial_learning_on_pointclouds_b200.trainer import adversarial_cls_step
    import argparse
    torch.manual_seed(4)
    g = M.PointNetCls(40, False)
    d = init_net(M.DeepConvDiscNet(40, 1), "cpu", "xavier")
    randomize_biases([g, d], 8)
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    g.to(DEV); d.to(DEV)
    g.dropout.p = 0.0
    for mod in list(g.modules()) + list(d.modules()):
        mod.precision = Precision(mode)
    pts, y, _, _ = inputs(6, 200, 31)
    pts2, _, _, _ = inputs(6, 200, 32)
    opt = torch.optim.SGD(g.parameters(), lr=0.0)
    optD = torch.optim.SGD(d.parameters(), lr=0.0)
    targs = argparse.Namespace(device=DEV, lambda_cls=1.0, lambda_adv=0.5)
    torch.manual_seed(77)
    l_cls, l_adv, l_D = adversarial_cls_step(g, d, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(),
                                             opt, optD, (pts.to(DEV), y.to(DEV)), pts2.to(DEV), targs)
    torch.manual_seed(77)
    ref = steps.adversarial_cls_step(gp, dp, (pts, y), (pts2,), lambda_adv=0.5)
    tol = TOL[mode]
    assert abs(l_cls.item() - ref["l_cls"]) < 10 * tol and abs(l_adv.item() - ref["l_adv"]) < 10 * tol
    assert abs(l_D.item() - (ref["l_D_gt"] + ref["l_D_nogt"])) < 10 * tol
    loose = 5e-3 if mode == "fp32" else 0.15                                  # un-conditioned, as above
    for k, v in d.named_parameters():
        assert rel_err(v.grad, dp[k].grad) < loose, k
    for k, v in g.named_parameters():
        assert rel_err(v.grad, gp[k].grad) < loose, k


# ------------------------------------------------------------- fused loss heads
@pytest.mark.parametrize("out_dtype,cols", [(torch.float32, 50), (torch.float16, 64), (torch.bfloat16, 64)])
@pytest.mark.parametrize("rows", [1, 77, 4096 + 5])
def test_softmax_head_kernels(out_dtype, cols, rows):
    """pcadv_softmax_head / pcadv_logsoftmax_bwd against torch (fp64) on the same logits."""
    n = 50
    logits = (_rand((rows, n), 5) * 3).to(DEV)
    labels = torch.randint(0, n, (rows,), generator=torch.Generator().manual_seed(6)).to(DEV)
    ref = logits.double()
    tol = 2e-6 if out_dtype == torch.float32 else (2e-3 if out_dtype == torch.float16 else 1.6e-2)
    loss_sum = torch.zeros(1, device=DEV)
    probs, dz = ops.softmax_head(logits, ops.HEAD_CE, labels=labels, out_dtype=out_dtype, cols=cols,
                                 want_dz=True, dz_gain=4.0, loss_sum=loss_sum)
    sm = torch.softmax(ref, 1)
    assert rel_err(probs[:, :n], sm) < tol
    assert rel_err(dz[:, :n], 4.0 * (sm - F.one_hot(labels, n).double())) < tol
    assert probs[:, n:].abs().max().item() == 0 if cols > n else True
    assert dz[:, n:].abs().max().item() == 0 if cols > n else True
    ce = F.cross_entropy(ref, labels, reduction="sum").item()
    assert abs(loss_sum.item() - ce) <= 2e-6 * abs(ce) + 1e-6
    lp, none = ops.softmax_head(logits, ops.HEAD_LSM, out_dtype=out_dtype, cols=cols)
    assert none is None
    assert rel_err(lp[:, :n], torch.log_softmax(ref, 1)) < tol
    # log_softmax backward from the saved (rounded) lp and a random incoming gradient
    dy = _rand((rows, cols), 7).to(DEV).to(out_dtype)
    dy[:, n:] = 0
    sc = torch.tensor([0.5], device=DEV)
    got = ops.logsoftmax_bwd(lp, dy, n, scale=sc, out_dtype=out_dtype, cols=cols)
    lpd, dyd = lp[:, :n].double(), dy[:, :n].double()
    want = 0.5 * (dyd - lpd.exp() * dyd.sum(1, keepdim=True))
    assert rel_err(got[:, :n], want) < tol
    if cols > n:
        assert got[:, n:].abs().max().item() == 0


@pytest.mark.parametrize("mode", MODES)
def test_fused_adversarial_step_vs_oracle_step(mode):
    """trainer.adversarial_seg_step_fused (fused CE / softmax / log_softmax heads, packed
    16-bit discriminator inputs) against oracle.steps: same losses, same gradients."""
    from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step_fused
    import argparse
    torch.manual_seed(1)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(384, 50), "cpu", "xavier")
    randomize_biases([g, d], 3)
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    g.to(DEV); d.to(DEV)
    g.precision = d.precision = Precision(mode)
    pts, _, seg, cls = inputs(3, 384, 8)
    pts2, _, _, cls2 = inputs(3, 384, 9)
    opt = torch.optim.SGD(g.parameters(), lr=0.0)
    optD = torch.optim.SGD(d.parameters(), lr=0.0)
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=0.5)
    torch.manual_seed(77)
    l_seg, l_adv, l_D = adversarial_seg_step_fused(
        g, d, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), opt, optD,
        tuple(t.to(DEV) for t in (pts, cls, seg)), tuple(t.to(DEV) for t in (pts2, cls2)), targs)
    torch.manual_seed(77)
    ref = steps.adversarial_seg_step(gp, dp, (pts, cls, seg), (pts2, cls2), lambda_adv=0.5)
    tol = TOL[mode]
    assert abs(l_seg.item() - ref["l_seg"]) < 10 * tol and abs(l_adv.item() - ref["l_adv"]) < 10 * tol
    loose = max(50 * tol, 5e-3)
    for k, v in d.named_parameters():
        assert rel_err(v.grad, dp[k].grad) < loose, k
    for k, v in g.named_parameters():
        assert rel_err(v.grad, gp[k].grad) < loose, k


@pytest.mark.parametrize("optim", ["sgd", "adam"])
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("mode", MODES)
def test_training_trajectory_tracks_oracle(mode, fused, optim):
    """30 full iterations (forward, backward, optimizer step on G and D, fresh batch and fresh
    smoothed labels each time) next to the oracle doing the same on the CPU.

    With SGD the deviation accumulates smoothly and is bounded directly.  Adam (what the reference
    trains with) turns rounding-level gradient differences into lr-sized parameter differences, so
    two correct fp32 implementations drift apart too -- and this path's atomics make the drift vary
    from run to run.  The yardstick there is the drift of the SAME oracle run by stock PyTorch eager
    on the GPU (TF32 off), which is itself several per cent on the losses after 30 iterations (the
    discriminator's max over channels and the max-pool switch branches).  Parameter distances are
    relative to how far training moved the parameters."""
    from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step, adversarial_seg_step_fused
    import argparse
    iters, B, N = 30, 3, 384
    torch.manual_seed(5)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
    randomize_biases([g, d], 4)
    start = {k: v.clone() for k, v in g.state_dict().items()}
    if optim == "adam":                                                      # 10 x train_segmentation.py:134-146
        mk = lambda ps, lr: torch.optim.Adam(ps, lr=lr, betas=(0.9, 0.999))
    else:
        mk = lambda ps, lr: torch.optim.SGD(ps, lr=50 * lr)
    arms = {}
    for name, dev in (("cpu", "cpu"), ("eager", DEV)):                      # the oracle on two ATen backends
        gp = steps.leaf_params({k: v.to(dev) for k, v in g.state_dict().items()})
        dp = steps.leaf_params({k: v.to(dev) for k, v in d.state_dict().items()})
        arms[name] = (gp, dp, mk(list(gp.values()), 1e-3), mk(list(dp.values()), 1e-4), dev)
    g.to(DEV); d.to(DEV)
    g.precision = d.precision = Precision(mode)
    opt, optD = mk(g.parameters(), 1e-3), mk(d.parameters(), 1e-4)
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=1e-3)
    step_fn = adversarial_seg_step_fused if fused else adversarial_seg_step
    old_tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    worst = {"ours": 0.0, "eager": 0.0}
    try:
        for it in range(iters):
            pts, _, seg, cls = inputs(B, N, 100 + it)
            pts2, _, _, cls2 = inputs(B, N, 500 + it)
            torch.manual_seed(1000 + it)                                     # the CPU label draws
            l_seg, l_adv, l_D = step_fn(g, d, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), opt,
                                        optD, tuple(t.to(DEV) for t in (pts, cls, seg)),
                                        tuple(t.to(DEV) for t in (pts2, cls2)), targs)
            got = {"ours": (l_seg.item(), l_adv.item(), l_D.item())}
            for name, (gp, dp, ropt, roptD, dev) in arms.items():
                torch.manual_seed(1000 + it)
                ropt.zero_grad(); roptD.zero_grad()
                r = steps.adversarial_seg_step(gp, dp, tuple(t.to(dev) for t in (pts, cls, seg)),
                                               tuple(t.to(dev) for t in (pts2, cls2)))
                ropt.step(); roptD.step()
                got[name] = (r["l_seg"], r["l_adv"], r["l_D_gt"] + r["l_D_nogt"])
            for name in worst:
                for mine, want in zip(got[name], got["cpu"]):
                    worst[name] = max(worst[name], abs(mine - want) / abs(want))
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_tf32
    ref = arms["cpu"][0]
    flat = lambda get: torch.cat([get(k).detach().cpu().flatten() for k in ref])
    theta = flat(lambda k: ref[k])
    moved = (theta - flat(lambda k: start[k])).norm()
    mine = dict(g.named_parameters())
    apart = {"ours": ((flat(lambda k: mine[k]) - theta).norm() / moved).item(),
             "eager": ((flat(lambda k: arms["eager"][0][k]) - theta).norm() / moved).item()}
    print("trajectory %s fused=%s %s: moved %.3f; worst loss deviation ours %.2e / eager-vs-cpu %.2e; "
          "parameters apart / moved: ours %.3e / eager-vs-cpu %.3e"
          % (mode, fused, optim, moved.item(), worst["ours"], worst["eager"], apart["ours"], apart["eager"]))
    assert moved > 0.03
    if optim == "sgd":
        # measured: 1.0-2.4 x the yardstick on the losses, 1.1-1.9 x on the parameters
        assert worst["ours"] < max(8 * worst["eager"], 2e-3)
        assert apart["ours"] < max(8 * apart["eager"], 4e-3)
    else:
        # measured: losses within 5e-3 (fp32) / 3e-2 (fp16) against 2e-3-4e-3 for the yardstick;
        # parameters 0.28-0.52 of the distance moved against 0.09-0.12.  The amplification is chaotic
        # (and this path's atomics make it vary from run to run), so these are sanity bounds: the
        # losses agree to a few per cent and the run does not wander off.
        assert worst["ours"] < 0.2
        assert apart["ours"] < 1.5


@pytest.mark.parametrize("mode", MODES)
def test_fused_heads_match_reference_shaped_forward(mode):
    """forward_ce / forward_logsoftmax against forward() + torch losses on the same module:
    loss, discriminator inputs and generator gradients."""
    g = build_seg(11, 12).to(DEV)
    g.precision = Precision(mode)
    pts, _, seg, cls = inputs(2, 300, 21)
    pts, seg, cls = pts.to(DEV), seg.to(DEV), cls.to(DEV)
    tol = TOL[mode]
    pred, _ = g(pts, cls)
    l_ref = F.cross_entropy(pred, seg)
    g.zero_grad(); (3.0 * l_ref).backward()
    ref_grads = {k: v.grad.clone() for k, v in g.named_parameters()}
    loss, probs, glob = g.forward_ce(pts, cls, seg)
    g.zero_grad(); (3.0 * loss).backward()
    assert abs(loss.item() - l_ref.item()) < 10 * tol
    sm = F.softmax(pred.detach(), dim=1)                                  # B x 50 x N
    got = probs[:, :, :50].transpose(1, 2) if probs.dtype != torch.float32 else probs
    assert rel_err(got, sm) < max(tol, 2e-3 if mode != "fp32" else tol)
    assert not probs.requires_grad
    for k, v in g.named_parameters():
        assert rel_err(v.grad, ref_grads[k]) < max(20 * tol, 2e-3 if mode != "fp32" else 0), k


def test_fused_adversarial_step_bf16():
    """The bf16 storage mode runs the same kernels (wide range, 8-bit mantissa): reported at ~1e-2,
    not claimed at 1e-3."""
    from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step_fused
    import argparse
    torch.manual_seed(1)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(384, 50), "cpu", "xavier")
    randomize_biases([g, d], 3)
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    g.to(DEV); d.to(DEV)
    g.precision = d.precision = Precision("bf16")
    pts, _, seg, cls = inputs(3, 384, 8)
    pts2, _, _, cls2 = inputs(3, 384, 9)
    opt = torch.optim.SGD(g.parameters(), lr=0.0)
    optD = torch.optim.SGD(d.parameters(), lr=0.0)
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=0.5)
    torch.manual_seed(77)
    l_seg, l_adv, _ = adversarial_seg_step_fused(
        g, d, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), opt, optD,
        tuple(t.to(DEV) for t in (pts, cls, seg)), tuple(t.to(DEV) for t in (pts2, cls2)), targs)
    torch.manual_seed(77)
    ref = steps.adversarial_seg_step(gp, dp, (pts, cls, seg), (pts2, cls2), lambda_adv=0.5)
    assert abs(l_seg.item() - ref["l_seg"]) < 2e-2 and abs(l_adv.item() - ref["l_adv"]) < 2e-2
    for k, v in list(g.named_parameters()) + list(d.named_parameters()):
        assert torch.isfinite(v.grad).all(), k
    for k, v in g.named_parameters():
        # un-conditioned comparison: with an 8-bit mantissa ~2^-8 of the ReLU / argmax decisions
        # land on the other side, which moves a gradient tensor by ~sqrt(2^-8) of its norm
        assert rel_err(v.grad, gp[k].grad) < 0.15, (k, rel_err(v.grad, gp[k].grad))


@pytest.mark.parametrize("B,N,k", [(3, 100, 3), (2, 257, 64), (4, 1000, 128), (1, 5, 7)])
def test_tnet_bmm_and_regulariser_kernels(B, N, k):
    """pcadv_bmm / pcadv_bmm_tgrad / pcadv_ortho_reg(_bwd) against torch autograd (fp64)."""
    from adversarial_learning_on_pointclouds_b200.models._mlp import BmmFunction, RegularizerFunction
    x = _rand((B, N, k), 41).to(DEV).requires_grad_(True)
    T = (_rand((B, k, k), 42) * 0.3 + torch.eye(k, device=DEV)).requires_grad_(True)
    y = BmmFunction.apply(x, T)
    reg = RegularizerFunction.apply(T)
    wy = _rand((B, N, k), 43).to(DEV)
    (3.0 * reg + (y * wy).sum()).backward()
    xd, Td = x.detach().double().requires_grad_(True), T.detach().double().requires_grad_(True)
    yd = torch.bmm(xd, Td)
    I = torch.eye(k, dtype=torch.float64, device=DEV)[None]
    regd = torch.mean(torch.norm(torch.bmm(Td, Td.transpose(2, 1)) - I, dim=(1, 2)))
    (3.0 * regd + (yd * wy.double()).sum()).backward()
    assert rel_err(y, yd) < 1e-5 and abs(reg.item() - regd.item()) < 1e-5 * max(1.0, abs(regd.item()))
    assert rel_err(x.grad, xd.grad) < 1e-5
    assert rel_err(T.grad, Td.grad) < 2e-5


@pytest.mark.parametrize("mode", MODES)
def test_empty_batch_and_single_point(mode):
    """Edge shapes: an empty batch launches nothing and yields empty outputs and zero gradients;
    clouds of one point work (the max over points is that point)."""
    g = build_seg(1, 2).to(DEV)
    d = init_net(M.PointwiseDiscNet(64, 50), "cpu", "xavier").to(DEV)
    g.precision = d.precision = Precision(mode)
    pts = torch.zeros(0, 64, 3, device=DEV)
    cls = torch.zeros(0, 1, 16, device=DEV)
    pred, glob = g(pts, cls)
    assert tuple(pred.shape) == (0, 50, 64) and tuple(glob.shape) == (0, 2048, 1)
    out = d(F.log_softmax(pred, 1))
    assert tuple(out.shape) == (0, 64)
    (pred.sum() + out.sum()).backward()
    assert all(p.grad is None or p.grad.abs().sum().item() == 0 for p in g.parameters())
    pts, _, seg, cls = inputs(2, 1, 5)
    pred, glob = g(pts.to(DEV), cls.to(DEV))
    gp = steps.leaf_params({k: v.detach().cpu() for k, v in g.state_dict().items()})
    o_pred, o_glob = PO.pointnet_seg_forward(gp, pts, cls)
    assert rel_err(pred, o_pred) < (1e-5 if mode == "fp32" else 2e-3)


def test_launch_counter_counts():
    before = pkg._lib.launch_count()
    x = _rand((256, 64), 1)
    ops.linear([x], _rand((64, 64), 2))
    assert pkg._lib.launch_count() == before + 1


# ------------------------------------------------------- PointNetSeg_regulization (a7)
@pytest.mark.parametrize("mode", MODES)
def test_seg_regulization_golden(golden, mode):
    G = golden["small_seg_regu"]
    net = build_seg(3, 11, regu=True).to(DEV)
    for mod in net.modules():
        mod.precision = Precision(mode)
    check_weights(net, G["weights"])
    pts, _, seg, cls = inputs(3, 200, 77)
    pred, glob, tf = net(pts.to(DEV), cls.to(DEV))
    reg = M.feature_transform_regularizer(tf)
    loss = F.cross_entropy(pred, seg.to(DEV)) + 0.5 * glob.square().mean() + 1e-3 * reg
    loss.backward()
    tol = TOL[mode]
    assert tuple(pred.shape) == (3, 50, 200) and tuple(tf.shape) == (3, 128, 128)
    assert abs(loss.item() - G["loss"]) < 10 * tol * abs(G["loss"])
    assert rel_err(pred, G["pred"]) < 4 * tol and rel_err(glob, G["glob"]) < 4 * tol
    if mode == "fp32":
        for k, v in net.named_parameters():
            assert_summary_close(v.grad, G["grads"][k], 3e-4, k)


# ------------------------------------ BASELINE full sizes: size-independent properties
def test_full_size_properties_cfg5():
    """cfg5 shape (B=256, N=4096 per pass, 2^20 points), fast mode.  The oracle cannot run
    this size, so the checks are properties of the path: a cloud's output does not depend on
    the rest of the batch (bit-exact), permuting a cloud's points permutes its logits and
    leaves the pooled feature unchanged, and the backward is linear in the incoming gradient."""
    B, N = 256, 4096
    net = build_seg(0).to(DEV)
    net.precision = Precision("fp16")
    pts, _, seg, cls = inputs(B, N, 1234)
    pts, cls, seg = pts.to(DEV), cls.to(DEV), seg.to(DEV)
    with torch.no_grad():
        pred, glob = net(pts, cls)
        sub, gsub = net(pts[37:41], cls[37:41])
        assert torch.equal(pred[37:41], sub) and torch.equal(glob[37:41], gsub)
        perm = torch.randperm(N, generator=torch.Generator().manual_seed(5)).to(DEV)
        p2, g2 = net(pts[:8][:, perm], cls[:8])
        assert torch.equal(g2, glob[:8])
        assert torch.equal(p2, pred[:8][:, :, perm])
    assert torch.isfinite(pred).all()
    # linearity of the backward in dL/dlogits: a power-of-two factor only shifts the dynamic
    # gradient scale, so the two runs differ by the order of the fp32 RED accumulation in wgrad
    small = (pts[:16], cls[:16])
    grads = []
    for factor in (1.0, 4.0):
        net.zero_grad()
        pr, _ = net(*small)
        (F.cross_entropy(pr, seg[:16]) * factor).backward()
        grads.append({k: v.grad.clone() for k, v in net.named_parameters()})
    for k in grads[0]:
        assert rel_err(grads[1][k], 4.0 * grads[0][k]) < 1e-4, k


def test_full_size_adversarial_step_runs_cfg5_shapes():
    """One adversarial iteration at the benchmark's per-GPU size (reduced to 64+64 clouds to
    keep the test short): finite losses, every parameter of G and D updated."""
    from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step
    import argparse
    B, N = 64, 4096
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier").to(DEV)
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier").to(DEV)
    before = {k: v.clone() for k, v in list(g.state_dict().items()) + list(d.state_dict().items())}
    opt = torch.optim.Adam(g.parameters(), lr=1e-4)
    optD = torch.optim.Adam(d.parameters(), lr=1e-5)
    pts, _, seg, cls = inputs(B, N, 1234)
    pts2, _, _, cls2 = inputs(B, N, 4321)
    args = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=1e-3)
    losses = adversarial_seg_step(g, d, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), opt,
                                  optD, (pts.to(DEV), cls.to(DEV), seg.to(DEV)),
                                  (pts2.to(DEV), cls2.to(DEV)), args)
    assert all(torch.isfinite(l) for l in losses)
    after = dict(list(g.state_dict().items()) + list(d.state_dict().items()))
    assert all(not torch.equal(before[k], after[k]) for k in before)

"""Branch-conditioned gradient parity for EVERY model class and for the full
adversarial step, in the benched precision (fp16 operands, fp32 accumulate: <= 1e-3 over all
gradients together, <= 4e-3 per tensor) and in the fp32 verification mode (1e-5 / 4e-5), at small
sizes and at the BASELINE.json sizes (cfg1 32 x 2500, cfg2 32 x 2500, cfg3 16 + 16 x 2048 full step,
cfg4 128 x 2048).

Method (tests/parity.py): the CUDA forward runs under a tape that keeps its ReLU / LeakyReLU signs
and max-pool argmaxes; the oracle runs once with its own decisions (forward parity; every decision
of the CUDA path that differs must sit within ``flip_bounds`` -- a few unit roundoffs of the format
-- of its boundary) and once with the CUDA path's decisions (loss and gradient parity)."""
import argparse

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from adversarial_learning_on_pointclouds_b200 import models as M, Precision                      # noqa: E402
from adversarial_learning_on_pointclouds_b200.models import _mlp                                   # noqa: E402
from adversarial_learning_on_pointclouds_b200.trainer import adversarial_seg_step_fused           # noqa: E402
from adversarial_learning_on_pointclouds_b200.utils import init_net                               # noqa: E402
from oracle import pointnet_oracle as PO, discriminator_oracle as DO, steps                       # noqa: E402
from helpers import build_seg, inputs, randomize_biases, rel_err                                   # noqa: E402
import parity                                                                                      # noqa: E402

DEV = "cuda"
TOL = {"fp32": 1e-5, "fp16": 1e-3}
MODES = ["fp32", "fp16"]


def _set_mode(model, mode):
    for mod in model.modules():
        mod.precision = Precision(mode)


def _cpu_sd(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def _report(tag, rep):
    print("%s: fwd %s loss %.2e grad_total %.2e worst tensor %.2e [stock TF32 eager, same decisions: fwd %s "
          "grad_total %.2e worst %.2e]; %d decisions differ, worst flipped margins: act %.2f u, argmax %.2f u"
          % (tag, ["%.1e" % e for e in rep["fwd"]], rep["loss"], rep["grad_total"], max(rep["grads"].values()),
             ["%.1e" % e for e in rep["yard_fwd"]], rep["yard_total"],
             max(rep["yard_grads"].values()) if rep["yard_grads"] else 0.0, rep["n_branch_diff"],
             rep["worst_act_margin_u"], rep["worst_argmax_margin_u"]))


_FEAT = dict(layers=[("feat.conv1", "bcn"), ("feat.conv2", "bcn"), ("feat.conv3", "bcn")],
             reduce=("feat.conv4", "points"))


def _stn_plan(tag):
    return [dict(layers=[(tag + ".conv1", "bcn"), (tag + ".conv2", "bcn")], reduce=(tag + ".conv3", "points")),
            dict(layers=[(tag + ".fc1", "bc"), (tag + ".fc2", "bc"), None])]


# ------------------------------------------------------------------------------- PointNetCls (a3, a5)
def _cls_parity(mode, ft, B, N, seed):
    torch.manual_seed(seed)
    m = M.PointNetCls(40, ft)
    randomize_biases([m], seed + 1)
    sd = _cpu_sd(m)
    m.to(DEV).eval()
    _set_mode(m, mode)
    pts, y, _, _ = inputs(B, N, seed + 2)
    if ft:
        plan = [dict(layers=[("feat.conv1", "bcn"), ("feat.conv2", "bcn")])] + _stn_plan("feat.fstn") + \
               [dict(layers=[("feat.conv3", "bcn")], reduce=("feat.conv4", "points"))]
    else:
        plan = [_FEAT]
    plan = plan + [dict(layers=[("fc1", "bc"), ("fc2", "bc"), None])]

    def cuda_run():
        logits, glob, tf = m(pts.to(DEV))
        loss = F.cross_entropy(logits, y.to(DEV)) + 0.5 * glob.square().mean()
        if ft:
            loss = loss + 1e-3 * M.feature_transform_regularizer(tf)
        return [logits, glob] + ([tf] if ft else []), loss

    def oracle_run(params, branch, record, dev):
        logits, glob, tf = PO.pointnet_cls_forward(params, pts.to(dev), ft, branch=branch, record=record)
        loss = F.cross_entropy(logits, y.to(dev)) + 0.5 * glob.square().mean()
        if ft:
            loss = loss + 1e-3 * PO.feature_transform_regularizer(tf)
        return [logits, glob] + ([tf] if ft else []), loss

    return parity.branch_parity(mode, TOL[mode], cuda_run, oracle_run, plan, list(m.named_parameters()), sd, B)


@pytest.mark.parametrize("ft", [False, True])
@pytest.mark.parametrize("mode", MODES)
def test_cls_branch_parity(mode, ft):
    _report("PointNetCls ft=%s %s" % (ft, mode), _cls_parity(mode, ft, 4, 600, 40))


# ------------------------------------------------------------------------------- PointNetDenseCls (a4)
def _dense_parity(mode, B, N, seed):
    torch.manual_seed(seed)
    m = M.PointNetDenseCls(num_classes=50)
    randomize_biases([m], seed + 1)
    sd = _cpu_sd(m)
    m.to(DEV)
    _set_mode(m, mode)
    pts, _, seg, _ = inputs(B, N, seed + 2)
    x = pts.transpose(1, 2).contiguous()
    plan = [_FEAT, dict(layers=[None]),
            dict(layers=[("conv1", "bcn"), ("conv2", "bcn"), ("conv3", "bcn"), None])]

    def cuda_run():
        out, _ = m(x.to(DEV))
        return [out], F.nll_loss(out.reshape(-1, 50), seg.reshape(-1).to(DEV))

    def oracle_run(params, branch, record, dev):
        out, _ = PO.pointnet_densecls_forward(params, x.to(dev), 50, branch=branch, record=record)
        return [out], F.nll_loss(out.reshape(-1, 50), seg.reshape(-1).to(dev))

    return parity.branch_parity(mode, TOL[mode], cuda_run, oracle_run, plan, list(m.named_parameters()), sd, B)


@pytest.mark.parametrize("mode", MODES)
def test_densecls_branch_parity(mode):
    _report("PointNetDenseCls %s" % mode, _dense_parity(mode, 3, 500, 50))


# ------------------------------------------------------------------------- PointNetSeg_regulization (a7)
@pytest.mark.parametrize("mode", MODES)
def test_seg_regulization_branch_parity(mode):
    B, N = 3, 400
    m = build_seg(3, 11, regu=True)
    sd = _cpu_sd(m)
    m.to(DEV)
    _set_mode(m, mode)
    pts, _, seg, cls = inputs(B, N, 77)
    one = lambda name, lay="bcn": dict(layers=[(name, lay)])
    plan = _stn_plan("stn") + [one("conv1"), one("conv2"), one("conv3")] + _stn_plan("fstn") + \
        [one("conv4"), one("conv5"), dict(layers=[], reduce=("conv6", "points")), dict(layers=[None]),
         dict(layers=[("fc1", "bnc"), ("fc2", "bnc"), ("fc3", "bnc"), None])]

    def cuda_run():
        pred, glob, tf = m(pts.to(DEV), cls.to(DEV))
        loss = F.cross_entropy(pred, seg.to(DEV)) + 0.5 * glob.square().mean() + \
            1e-3 * M.feature_transform_regularizer(tf)
        return [pred, glob, tf], loss

    def oracle_run(params, branch, record, dev):
        pred, glob, tf = PO.pointnet_seg_forward(params, pts.to(dev), cls.to(dev), regulization=True, branch=branch,
                                                 record=record)
        loss = F.cross_entropy(pred, seg.to(dev)) + 0.5 * glob.square().mean() + \
            1e-3 * PO.feature_transform_regularizer(tf)
        return [pred, glob, tf], loss

    _report("PointNetSeg_regulization %s" % mode,
            parity.branch_parity(mode, TOL[mode], cuda_run, oracle_run, plan, list(m.named_parameters()), sd, B))


# ------------------------------------------------------------------------------ discriminators (a9-a13)
def _disc_case(name, N):
    """-> (modules, plan, cuda forward(mods, x), oracle forward(params, x, branch, record))."""
    chain = lambda names, lay="bcn": [(n, lay) for n in names]
    if name == "pointwise":
        mods = [M.PointwiseDiscNet(N, 50)]
        plan = [dict(layers=chain(["conv1", "conv2", "conv3"]), reduce=("conv4", "channels"))]
        return mods, plan, (lambda m, x: [m[0](x)]), \
            (lambda p, x, b, r: [DO.pointwise_disc_forward(p[0], x, N, branch=b, record=r)])
    if name == "conv":
        mods = [M.ConvDiscNet(50)]
        plan = [dict(layers=chain(["conv1", "conv2", "conv3"]) + [None])]
        return mods, plan, (lambda m, x: [m[0](x.transpose(1, 2))]), \
            (lambda p, x, b, r: [DO.conv_disc_forward(p[0], x.transpose(1, 2), branch=b, record=r)])
    if name == "stack":
        mods = [M.StackDiscNet(N, 50, 16)]
        plan = [dict(layers=chain(["conv1", "conv2", "conv3"]), reduce=("conv4", "channels")), dict(layers=[None])]
        return mods, plan, (lambda m, x: list(m[0](x))), \
            (lambda p, x, b, r: list(DO.stack_disc_forward(p[0], x, branch=b, record=r)))
    if name == "dual":
        mods = [M.BaseDiscNet(N, 50, 256), M.ShapeDiscNet(256, 16), M.PointDiscNet(256, N)]
        plan = [dict(layers=chain(["base.conv1", "base.conv2", "base.conv3"])),
                dict(layers=[], reduce=("shape.conv", "points")), dict(layers=[("shape.fc1", "bc"), None]),
                dict(layers=chain(["point.conv1", "point.conv2"]), reduce=("point.conv3", "channels"))]

        def cuda_fwd(m, x):
            shared = m[0](x)
            return [m[1](shared), m[2](shared)]

        def oracle_fwd(p, x, b, r):
            shared = DO.base_disc_forward(p[0], x, branch=b, record=r)
            return [DO.shape_disc_forward(p[1], shared, branch=b, record=r),
                    DO.point_disc_forward(p[2], shared, N, branch=b, record=r)]
        return mods, plan, cuda_fwd, oracle_fwd
    raise ValueError(name)


@pytest.mark.parametrize("name", ["pointwise", "conv", "stack", "dual"])
@pytest.mark.parametrize("mode", MODES)
def test_discriminator_branch_parity(mode, name):
    B, N = 3, 640
    torch.manual_seed(31)
    mods, plan, cuda_fwd, oracle_fwd = _disc_case(name, N)
    mods = [init_net(mm, "cpu", "xavier") for mm in mods]
    randomize_biases(mods, 13)
    sds = [_cpu_sd(mm) for mm in mods]
    for mm in mods:
        mm.to(DEV)
        _set_mode(mm, mode)
    x0 = torch.log_softmax(torch.randn(B, 50, N, generator=torch.Generator().manual_seed(21)), dim=1)
    x = x0.clone().to(DEV).requires_grad_(True)
    weight = lambda o: torch.linspace(0.5, 1.5, o.numel()).view_as(o)
    flat_sd = {"%d.%s" % (i, k): v for i, sd in enumerate(sds) for k, v in sd.items()}
    named = [("%d.%s" % (i, k), v) for i, mm in enumerate(mods) for k, v in mm.named_parameters()]
    holder = {}

    def cuda_run():
        outs = cuda_fwd(mods, x)
        return outs, sum((o * weight(o).to(DEV)).mean() for o in outs)

    def oracle_run(params, branch, record, dev):
        per = [{k.split(".", 1)[1]: v for k, v in params.items() if k.startswith("%d." % i)} for i in range(len(mods))]
        xo = x0.clone().to(dev).requires_grad_(True)
        if dev == "cpu":
            holder["x"] = xo
        outs = oracle_fwd(per, xo, branch, record)
        return outs, sum((o * weight(o).to(dev)).mean() for o in outs)

    rep = parity.branch_parity(mode, TOL[mode], cuda_run, oracle_run, plan, named, flat_sd, B,
                               extra_grads=lambda _: [("dx", x.grad, holder["x"].grad)])
    _report("disc %s %s" % (name, mode), rep)


@pytest.mark.parametrize("mode", MODES)
def test_deepconv_discriminator_branch_parity(mode):
    B = 6
    torch.manual_seed(35)
    dd = init_net(M.DeepConvDiscNet(40, 1), "cpu", "xavier")
    randomize_biases([dd], 14)
    sd = _cpu_sd(dd)
    dd.to(DEV)
    _set_mode(dd, mode)
    x0 = torch.log_softmax(torch.randn(B, 40, generator=torch.Generator().manual_seed(22)), 1)
    x = x0.clone().to(DEV).requires_grad_(True)
    plan = [dict(layers=[("conv%d" % i, "bc1") for i in range(1, 6)] + [None])]
    holder = {}

    def cuda_run():
        o = dd(x)
        return [o], (o * torch.linspace(0.5, 1.5, o.numel(), device=DEV).view_as(o)).mean()

    def oracle_run(params, branch, record, dev):
        xo = x0.clone().to(dev).requires_grad_(True)
        if dev == "cpu":
            holder["x"] = xo
        o = DO.deepconv_disc_forward(params, xo, branch=branch, record=record)
        return [o], (o * torch.linspace(0.5, 1.5, o.numel(), device=dev).view_as(o)).mean()

    rep = parity.branch_parity(mode, TOL[mode], cuda_run, oracle_run, plan, list(dd.named_parameters()), sd, B,
                               extra_grads=lambda _: [("dx", x.grad, holder["x"].grad)])
    _report("disc deepconv %s" % mode, rep)


# ----------------------------------------------------------------- the full adversarial step (a15)
class _SegTape(list):
    """Stands in for the ``_debug`` dict of PointNetSeg: one record per SegFunction forward."""

    def update(self, **kw):
        self.append(kw)


def _split_debug(dbg, b0, b1, N):
    """The rows of clouds [b0, b1) of a SegFunction debug record."""
    sl = slice(b0 * N, b1 * N)
    return dict(x=[t[sl] for t in dbg["x"]], h=[t[sl] for t in dbg["h"]], idx=dbg["idx"][b0:b1])


def _adv_step_parity(mode, B, N, one_pass, lambda_adv, seed=1, in_seeds=(1234, 4321)):
    torch.manual_seed(seed)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
    randomize_biases([g, d], 3)
    g_sd, d_sd = _cpu_sd(g), _cpu_sd(d)
    g.to(DEV); d.to(DEV)
    g.precision = d.precision = Precision(mode)
    pts, _, seg, cls = inputs(B, N, in_seeds[0])
    pts2, _, _, cls2 = inputs(B, N, in_seeds[1])
    lg = torch.Generator().manual_seed(99)
    lab_r = torch.empty(B, N).uniform_(0.7, 1.05, generator=lg)
    lab_f = torch.empty(B, N).uniform_(0.0, 0.305, generator=lg)
    lab_dev = (lab_r.to(DEV), lab_f.to(DEV))
    label_fn = lambda d_out, value, rnd: (torch.full_like(d_out, float(value)) if not rnd
                                          else (lab_dev[0] if value == 1 else lab_dev[1]))
    opt, optD = torch.optim.SGD(g.parameters(), lr=0.0), torch.optim.SGD(d.parameters(), lr=0.0)
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=lambda_adv)
    g._debug = seg_tape = _SegTape()
    _mlp.DEBUG_TAPE = d_tape = []
    try:
        l_seg, l_adv, l_D = adversarial_seg_step_fused(
            g, d, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), opt, optD,
            tuple(t.to(DEV) for t in (pts, cls, seg)), tuple(t.to(DEV) for t in (pts2, cls2)), targs,
            label_fn=label_fn, one_pass=one_pass)
    finally:
        g._debug = None
        _mlp.DEBUG_TAPE = None
    if one_pass:
        assert len(seg_tape) == 1
        dbg_gt, dbg_nogt = _split_debug(seg_tape[0], 0, B, N), _split_debug(seg_tape[0], B, 2 * B, N)
    else:
        dbg_gt, dbg_nogt = seg_tape
    assert len(d_tape) == 3                       # D(lsm nogt) [G phase], D(softmax gt), D(lsm nogt) [cache hit]
    # pass 1: the oracle with its own decisions -> records; boundaries of every differing decision
    rec = {}
    gp, dp = steps.leaf_params(g_sd), steps.leaf_params(d_sd)
    steps.adversarial_seg_step(gp, dp, (pts, cls, seg), (pts2, cls2), lambda_adv=lambda_adv,
                               labels=(lab_r, lab_f), record=rec)
    d_plan = [dict(layers=[("conv1", "bcn"), ("conv2", "bcn"), ("conv3", "bcn")], reduce=("conv4", "channels"))]
    branch = {"g_gt": parity.seg_branches(dbg_gt, B, N), "g_nogt": parity.seg_branches(dbg_nogt, B, N)}
    for key, t in zip(("d_adv", "d_gt", "d_nogt"), d_tape):
        branch[key] = parity.tape_branches([t], d_plan, {k: v.detach() for k, v in rec[key].items()}, B)
    act_tol, arg_tol = parity.flip_bounds(mode)
    stats, n_diff = {}, 0
    for key in branch:
        n_diff += parity.check_branch_boundaries(branch[key], {k: v.detach() for k, v in rec[key].items()},
                                                 act_tol, arg_tol, stats)
    # pass 2: the oracle with the CUDA path's decisions -> losses and gradients
    gp, dp = steps.leaf_params(g_sd), steps.leaf_params(d_sd)
    ref = steps.adversarial_seg_step(gp, dp, (pts, cls, seg), (pts2, cls2), lambda_adv=lambda_adv,
                                     labels=(lab_r, lab_f), branch=branch)
    tol = TOL[mode]
    assert abs(l_seg.item() - ref["l_seg"]) <= tol * abs(ref["l_seg"])
    assert abs(l_adv.item() - ref["l_adv"]) <= tol * abs(ref["l_adv"])
    assert abs(l_D.item() - (ref["l_D_gt"] + ref["l_D_nogt"])) <= tol * abs(ref["l_D_gt"] + ref["l_D_nogt"])
    g_errs, g_tot = parity.grad_report(list(g.named_parameters()), gp)
    d_errs, d_tot = parity.grad_report(list(d.named_parameters()), dp)
    # yardstick (16-bit modes): the same oracle step with the same pinned decisions, run by stock
    # TF32 eager on the GPU -- what the arithmetic the north star names does to these gradients
    yg_errs, yg_tot, yd_errs, yd_tot = {}, 0.0, {}, 0.0
    if mode != "fp32":
        cu = lambda obj: parity._to(obj, DEV)
        ygp, ydp = steps.leaf_params(cu(g_sd)), steps.leaf_params(cu(d_sd))
        with parity.tf32_eager():
            steps.adversarial_seg_step(ygp, ydp, cu((pts, cls, seg)), cu((pts2, cls2)), lambda_adv=lambda_adv,
                                       labels=cu((lab_r, lab_f)), branch=cu(branch))
        yg_errs, yg_tot = parity.grad_report(list(ygp.items()), gp)
        yd_errs, yd_tot = parity.grad_report(list(ydp.items()), dp)
    print("adversarial step %s B=%d N=%d one_pass=%s lambda_adv=%g: G grads total %.2e worst %.2e; D grads total "
          "%.2e worst %.2e [stock TF32 eager, same decisions: G %.2e / %.2e, D %.2e / %.2e]; %d decisions differ, "
          "worst flipped margins: act %.2f u, argmax %.2f u"
          % (mode, B, N, one_pass, lambda_adv, g_tot, max(g_errs.values()), d_tot, max(d_errs.values()),
             yg_tot, max(yg_errs.values()) if yg_errs else 0.0, yd_tot, max(yd_errs.values()) if yd_errs else 0.0,
             n_diff, stats["worst_act_margin"] / parity.U[mode], stats["worst_argmax_margin"] / parity.U[mode]))
    parity.gate(g_tot, yg_tot, tol, "all G gradients")
    parity.gate(d_tot, yd_tot, tol, "all D gradients")
    for k, e in g_errs.items():
        parity.gate(e, yg_errs.get(k, 0.0), 4 * tol, "G." + k)
    for k, e in d_errs.items():
        parity.gate(e, yd_errs.get(k, 0.0), 4 * tol, "D." + k)


@pytest.mark.parametrize("lambda_adv", [1e-3, 0.5])
@pytest.mark.parametrize("one_pass", [True, False])
@pytest.mark.parametrize("mode", MODES)
def test_adversarial_step_branch_parity(mode, one_pass, lambda_adv):
    _adv_step_parity(mode, 3, 512, one_pass, lambda_adv)


# ------------------------------------------------------------------- BASELINE.json sizes (cfg1 - cfg4)
def test_cfg1_size_cls_branch_parity():
    """cfg1: PointNetCls(40), B = 32, N = 2500 (models/pointnet.py:186-203, :356), benched precision."""
    _report("cfg1 PointNetCls 32 x 2500 fp16", _cls_parity("fp16", False, 32, 2500, 0))


def test_cfg2_size_densecls_branch_parity():
    """cfg2: PointNetDenseCls(50), B = 32, N = 2500 (models/pointnet.py:320-343)."""
    _report("cfg2 PointNetDenseCls 32 x 2500 fp16", _dense_parity("fp16", 32, 2500, 6))


def test_cfg4_size_cls_feature_transform_branch_parity():
    """cfg4: PointNetCls(40, feature_transform=True) + regulariser, B = 128, N = 2048
    (models/pointnet.py:46-79, :109-136, :345-353)."""
    _report("cfg4 PointNetCls ft 128 x 2048 fp16", _cls_parity("fp16", True, 128, 2048, 5))


@pytest.mark.parametrize("mode", MODES)
def test_cfg3_size_adversarial_step_branch_parity(mode):
    """cfg3: the full adversarial step at 16 + 16 clouds of 2048 points (utils/trainer.py:873-966)."""
    _adv_step_parity(mode, 16, 2048, True, 1e-3, seed=0)

"""GPU: the remaining live loop bodies of utils/trainer.py on the CUDA modules --
``adversarial_seg_dual_step`` (:2150-2284), ``adversarial_cls_semi_step`` (:635-794),
``adversarial_seg_semi_step`` (:1927-2061) -- against ``oracle.steps`` (itself pinned to iterations
of the reference's unmodified functions, tests/test_oracle_golden.py) and, in the fp32 mode, directly
against those golden parameters.  Plus: the reference's own ``utils/trainer.py`` driving the new
modules unchanged (only where a checkout of the reference is reachable: ``PCADV_REFERENCE``)."""
import argparse
import logging
import os
import sys
import tempfile
import types

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from adversarial_learning_on_pointclouds_b200 import models as M, Precision                      # noqa: E402
from adversarial_learning_on_pointclouds_b200 import trainer as T                                  # noqa: E402
from adversarial_learning_on_pointclouds_b200.utils import init_net, ImagePool                    # noqa: E402
from oracle import steps                                                                           # noqa: E402
from helpers import assert_summary_close, inputs                                                   # noqa: E402
from test_oracle_golden import dual_setup                                                          # noqa: E402

DEV = "cuda"
MODES = ["fp32", "fp16"]


@pytest.fixture(scope="module")
def golden_steps():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_steps.pt"), weights_only=False)


def _set_mode(mods, mode):
    for m in mods:
        for sub in m.modules():
            sub.precision = Precision(mode)


def _distance(mod, params, start):
    """(|theta_cuda - theta_oracle|, |theta_oracle - theta_start|) over all parameters of a module."""
    apart = torch.cat([(v.detach().cpu() - params[k].detach()).flatten() for k, v in mod.named_parameters()]).norm()
    moved = torch.cat([(params[k].detach() - start[k]).flatten() for k, _ in mod.named_parameters()]).norm()
    return apart.item(), moved.item()


def _check(tag, mode, mods_params_start, golden=None):
    for name, mod, params, start in mods_params_start:
        apart, moved = _distance(mod, params, start)
        print("%s %s %s: parameters moved %.3e, CUDA path apart from the oracle %.3e" % (tag, mode, name, moved, apart))
        # fp32: summation order only, measured 1e-5 (one flipped decision under Adam: ~2.5e-3 of the distance moved);
        # fp16: the rounding of the mode under the sign-like first steps of Adam on 512-point batches (measured 0.14-0.17)
        assert apart <= (1e-2 if mode == "fp32" else 0.3) * moved + 1e-7, (name, apart, moved)
        if golden is not None and mode == "fp32":
            for k, v in mod.state_dict().items():
                # biases start at zero and are lr-sized after a few steps (1e-5 .. 1e-4), so one decision
                # that lands on the other side (atomics order, seen once in three runs) shows at ~1e-3 of
                # them: their probes are compared at 2e-3, weights at 5e-5.  The strict comparison is
                # oracle-vs-golden on the CPU (1e-6) plus the distance check above.
                assert_summary_close(v, golden[name][k], 2e-3 if k.endswith("bias") else 5e-5, "%s:%s" % (name, k))


@pytest.mark.parametrize("mode", MODES)
def test_adversarial_seg_dual_step(golden_steps, mode):
    G = golden_steps["dual"]
    R = G["recipe"]
    g, shared, shape, point = dual_setup(R["N"])
    mods = (g, shared, shape, point)
    start = [{k: v.clone() for k, v in m.state_dict().items()} for m in mods]
    gp, sp, hp, pp = (steps.leaf_params(m.state_dict()) for m in mods)
    ropt = torch.optim.Adam(list(gp.values()), lr=R["lr_g"], betas=(0.9, 0.999))
    ropt_shape = torch.optim.SGD(list(hp.values()) + list(sp.values()), lr=R["lr_d"])
    ropt_point = torch.optim.SGD(list(pp.values()) + list(sp.values()), lr=R["lr_d"])
    for m in mods:
        m.to(DEV)
    _set_mode(mods, mode)
    opt = torch.optim.Adam(g.parameters(), lr=R["lr_g"], betas=(0.9, 0.999))
    opt_shape = torch.optim.SGD(list(shape.parameters()) + list(shared.parameters()), lr=R["lr_d"])
    opt_point = torch.optim.SGD(list(point.parameters()) + list(shared.parameters()), lr=R["lr_d"])
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=R["lambda_adv"], lambda_disc_shape=1.0)
    gan, ce = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()
    batches = []
    for it in range(R["iters"]):
        pts, _, seg, cls = inputs(R["B"], R["N"], R["seed"] + it)
        pts2, _, _, cls2 = inputs(R["B"], R["N"], R["seed2"] + it)
        batches.append(((pts, cls, seg), (pts2, cls2)))
    torch.manual_seed(R["label_seed"])
    for bg, bn in batches:
        out = T.adversarial_seg_dual_step(g, shared, shape, point, gan, ce, ce, opt, opt_shape, opt_point,
                                          tuple(t.to(DEV) for t in bg), tuple(t.to(DEV) for t in bn), targs)
        assert all(torch.isfinite(o) for o in out)
    # the never-zeroed optimizer_D_point: pointDisc's gradient holds the sum over both iterations
    torch.manual_seed(R["label_seed"])
    for bg, bn in batches:
        ref = steps.adversarial_seg_dual_step(gp, sp, hp, pp, bg, bn, ropt, ropt_shape, ropt_point,
                                              lambda_adv=R["lambda_adv"])
    assert abs(out[0].item() - ref["l_seg"]) < (1e-4 if mode == "fp32" else 5e-3)
    assert abs(out[3].item() - ref["l_D_shape"]) < (1e-4 if mode == "fp32" else 5e-3)
    _check("dual", mode, [("g", g, gp, start[0]), ("shared", shared, sp, start[1]), ("shape", shape, hp, start[2]),
                          ("point", point, pp, start[3])], golden=G)


@pytest.mark.parametrize("mode", MODES)
def test_adversarial_cls_semi_step(golden_steps, mode):
    G = golden_steps["cls_semi"]
    R = G["recipe"]
    torch.manual_seed(0)
    g = M.PointNetCls(40, False)
    g.dropout.p = 0.0
    d = init_net(M.DeepConvDiscNet(40, 1), "cpu", "xavier")
    start = [{k: v.clone() for k, v in m.state_dict().items()} for m in (g, d)]
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    ropt = torch.optim.Adam(list(gp.values()), lr=1e-4, betas=(0.9, 0.999))
    roptD = torch.optim.Adam(list(dp.values()), lr=1e-5, betas=(0.9, 0.999))
    g.to(DEV); d.to(DEV)
    _set_mode((g, d), mode)
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(d.parameters(), lr=1e-5, betas=(0.9, 0.999))
    targs = argparse.Namespace(device=DEV, lambda_cls=1.0, lambda_adv=1e-3, lambda_semi=1.0,
                               semi_start=R["semi_start"], semi_TH=R["semi_TH"])
    gan, ce, semi = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), torch.nn.CrossEntropyLoss(ignore_index=255)
    batches = [((inputs(R["B"], R["N"], R["seed"] + it)[0], inputs(R["B"], R["N"], R["seed"] + it)[1]),
                inputs(R["B"], R["N"], R["seed2"] + it)[0]) for it in range(R["iters"])]
    saw = False
    torch.manual_seed(R["label_seed"])
    for it, (bg, bn) in enumerate(batches):
        out = T.adversarial_cls_semi_step(g, d, gan, ce, semi, opt, optD, tuple(t.to(DEV) for t in bg), bn.to(DEV),
                                          targs, it)
        saw |= out[2] is not None
    assert saw
    torch.manual_seed(R["label_seed"])
    for it, (bg, bn) in enumerate(batches):
        ref = steps.adversarial_cls_semi_step(gp, dp, bg, (bn,), ropt, roptD, it, R["semi_start"], R["semi_TH"])
    if mode == "fp32":
        assert abs(out[2].item() - ref["l_semi"]) < 1e-4
    _check("cls_semi", mode, [("g", g, gp, start[0]), ("d", d, dp, start[1])], golden=G)


@pytest.mark.parametrize("mode", MODES)
def test_adversarial_seg_semi_step(golden_steps, mode):
    G = golden_steps["seg_semi"]
    R = G["recipe"]
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(R["N"], 50), "cpu", "xavier")
    start = [{k: v.clone() for k, v in m.state_dict().items()} for m in (g, d)]
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    ropt = torch.optim.Adam(list(gp.values()), lr=1e-4, betas=(0.9, 0.999))
    roptD = torch.optim.Adam(list(dp.values()), lr=1e-5, betas=(0.9, 0.999))
    g.to(DEV); d.to(DEV)
    _set_mode((g, d), mode)
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(d.parameters(), lr=1e-5, betas=(0.9, 0.999))
    targs = argparse.Namespace(device=DEV, lambda_seg=1.0, lambda_adv=1e-3, lambda_semi=1.0,
                               semi_start=R["semi_start"], semi_TH=R["semi_TH"])
    gan, ce, semi = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(), torch.nn.CrossEntropyLoss(ignore_index=255)
    batches = []
    for it in range(R["iters"]):
        pts, _, seg, cls = inputs(R["B"], R["N"], R["seed"] + it)
        pts2, _, _, cls2 = inputs(R["B"], R["N"], R["seed2"] + it)
        batches.append(((pts, cls, seg), (pts2, cls2)))
    saw = False
    torch.manual_seed(R["label_seed"])
    for it, (bg, bn) in enumerate(batches):
        out = T.adversarial_seg_semi_step(g, d, gan, ce, semi, opt, optD, tuple(t.to(DEV) for t in bg),
                                          tuple(t.to(DEV) for t in bn), targs, it)
        saw |= out[2] is not None
    assert saw
    torch.manual_seed(R["label_seed"])
    for it, (bg, bn) in enumerate(batches):
        ref = steps.adversarial_seg_semi_step(gp, dp, bg, bn, ropt, roptD, it, R["semi_start"], R["semi_TH"])
    if mode == "fp32":
        assert abs(out[2].item() - ref["l_semi"]) < 1e-3 * abs(ref["l_semi"])
    _check("seg_semi", mode, [("g", g, gp, start[0]), ("d", d, dp, start[1])], golden=G)


# --------------------------------------------------------------------------------------------------
def _reference_root():
    root = os.environ.get("PCADV_REFERENCE", "/root/reference")
    return root if os.path.exists(os.path.join(root, "utils", "trainer.py")) else None


@pytest.mark.skipif(_reference_root() is None, reason="no checkout of the reference reachable (PCADV_REFERENCE)")
def test_reference_trainer_drives_the_new_modules_unchanged():
    """The north star's "utils/trainer.py runs unchanged": import the reference's own trainer (with
    the matplotlib / np.object shims of SURVEY.md D6) and let its ``run_training_seg`` train the
    libpcadv-backed PointNetSeg / PointwiseDiscNet for one iteration, evaluation pass included."""
    import numpy as np
    root = _reference_root()
    if "matplotlib" not in sys.modules:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        plt.switch_backend = lambda *a, **k: None
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    for name, val in (("object", object), ("bool", bool), ("int", int)):
        if not hasattr(np, name):
            setattr(np, name, val)
    import importlib.util
    # the reference's package names (utils, models) collide with nothing of ours at top level
    sys.path.insert(0, root)
    try:
        spec = importlib.util.spec_from_file_location("ref_trainer", os.path.join(root, "utils", "trainer.py"))
        RT = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(RT)
    finally:
        sys.path.remove(root)
    N, B = 256, 2
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier").to(DEV)
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier").to(DEV)
    before = {k: v.clone() for k, v in list(g.state_dict().items()) + list(d.state_dict().items())}
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(d.parameters(), lr=1e-5, betas=(0.9, 0.999))
    pts, _, seg, cls = inputs(B, N, 1234)
    pts2, _, _, cls2 = inputs(B, N, 4321)
    tp, _, tseg, _ = inputs(16, N, 999)
    tcls = F.one_hot(torch.arange(16), 16).float().view(16, 1, 16)
    testloader = [(tp[i:i + 4], tcls[i:i + 4], tseg[i:i + 4]) for i in range(0, 16, 4)]
    a = argparse.Namespace(device=DEV, total_iterations=1, iter_save_epoch=1, iter_test_epoch=1, tensorboard=False,
                           exp_dir=tempfile.mkdtemp(), batch_size=B, input_pts=N, lambda_seg=1.0, lambda_adv=1e-3)
    logger = logging.getLogger("ref_trainer"); logger.setLevel("ERROR")
    RT.run_training_seg([(pts, cls, seg)], [(pts2, cls2)], enumerate([(pts, cls, seg)]), enumerate([(pts2, cls2)]),
                        testloader, list(range(16)), g, d, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(),
                        opt, optD, ImagePool(0), ImagePool(0), logger, logger, None, a)
    after = dict(list(g.state_dict().items()) + list(d.state_dict().items()))
    assert all(torch.isfinite(v).all() for v in after.values())
    assert all(not torch.equal(before[k], after[k]) for k in before)
    assert os.path.exists(os.path.join(a.exp_dir, "model_train_epoch_0.pth"))

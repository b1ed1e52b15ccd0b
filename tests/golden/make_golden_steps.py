#!/usr/bin/env python3
"""Golden fixtures for the remaining live loop bodies, produced by running the UNMODIFIED
reference trainer functions (imported read-only from /root/reference, CPU, with the shims of
SURVEY.md D6) for a few iterations on synthetic list loaders:

* ``run_training_seg_dual``  (utils/trainer.py:2123-2382) -- 2 iterations, so that the never-zeroed
  ``optimizer_D_point`` (:2171-2172) shows in the result;
* ``run_training_semi``      (utils/trainer.py:611-846)   -- 3 iterations with ``semi_start = 1``, so
  the third one adds the semi-supervised term (:727-739);
* ``run_training_seg_semi``  (utils/trainer.py:1903-2121) -- 3 iterations, same.

Run in the build container only (``python tests/golden/make_golden_steps.py``); writes
``tests/golden/golden_steps.pt``: the recipe plus summaries (norm, sum of |.|, seeded probes) of every
parameter after the last iteration.  ``tests/test_oracle_golden.py`` pins ``oracle/steps.py`` to it.
"""
import argparse
import logging
import os
import sys
import tempfile

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, inputs, install_shims, summarize        # noqa: E402


def pick_threshold(values):
    """A threshold in the widest gap between sorted values around the median: decisions
    ``value <= TH`` then do not depend on rounding."""
    v = torch.sort(values.flatten())[0]
    lo, hi = int(0.3 * len(v)), max(int(0.7 * len(v)), int(0.3 * len(v)) + 2)
    gaps = v[lo + 1:hi] - v[lo:hi - 1]
    i = int(torch.argmax(gaps)) + lo
    return float((v[i] + v[i + 1]) / 2), float(v[i + 1] - v[i])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(HERE, "golden_steps.pt"))
    args = ap.parse_args()
    install_shims()
    sys.path.insert(0, REF)
    torch.set_num_threads(8)
    from models.pointnet import PointNetSeg, PointNetCls
    from models import discriminator as RD
    from utils.model_utils import init_net
    from utils.image_pool import ImagePool
    from utils import trainer as RT
    logger = logging.getLogger("golden_steps"); logger.setLevel("ERROR")
    G = {}
    state = lambda m: {k: summarize(v) for k, v in m.state_dict().items()}

    # test loaders: all 16 categories / a classification test batch
    def seg_test(N):
        tp, _, tseg, _ = inputs(16, N, 999)
        tcls = F.one_hot(torch.arange(16), 16).float().view(16, 1, 16)
        return [(tp[i:i + 4], tcls[i:i + 4], tseg[i:i + 4]) for i in range(0, 16, 4)]

    # ---------------------------------------------------------------- run_training_seg_dual
    B, N, iters = 2, 256, 2
    torch.manual_seed(0)
    g = init_net(PointNetSeg(50), "cpu", "xavier")
    shared = init_net(RD.BaseDiscNet(N, 50, 256), "cpu", "xavier")
    shape = init_net(RD.ShapeDiscNet(256, 16), "cpu", "xavier")
    point = init_net(RD.PointDiscNet(256, N), "cpu", "xavier")
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999))          # train_segmentation.py:383-388
    opt_shape = torch.optim.SGD(list(shape.parameters()) + list(shared.parameters()), lr=1e-2)   # :395-401 (lr raised)
    opt_point = torch.optim.SGD(list(point.parameters()) + list(shared.parameters()), lr=1e-2)   # :396, :408-411
    gt = [tuple(t for t in (lambda p, y, s, c: (p, c, s))(*inputs(B, N, 1234 + i))) for i in range(iters)]
    nogt = [tuple(t for t in (lambda p, y, s, c: (p, c))(*inputs(B, N, 4321 + i))) for i in range(iters)]
    a = argparse.Namespace(device="cpu", total_iterations=iters, iter_save_epoch=1000, iter_test_epoch=1000,
                           tensorboard=False, exp_dir=tempfile.mkdtemp(), batch_size=B, input_pts=N,
                           lambda_seg=1.0, lambda_adv=1e-3, lambda_disc_shape=1.0)
    torch.manual_seed(4242)
    RT.run_training_seg_dual(gt, nogt, enumerate(gt), enumerate(nogt), seg_test(N), list(range(16)), g, shared,
                             shape, point, torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(),
                             torch.nn.CrossEntropyLoss(), opt, opt_shape, opt_point, ImagePool(0), ImagePool(0),
                             logger, logger, None, a)
    G["dual"] = dict(recipe=dict(B=B, N=N, iters=iters, seed=1234, seed2=4321, wseed=0, label_seed=4242,
                                 lr_g=1e-4, lr_d=1e-2, lambda_adv=1e-3),
                     g=state(g), shared=state(shared), shape=state(shape), point=state(point))
    print("dual done")

    # ---------------------------------------------------------------- run_training_semi (classification)
    B, N, iters = 6, 200, 3
    torch.manual_seed(0)
    g = PointNetCls(40, False)
    g.dropout.p = 0.0                       # the random stream of Dropout cannot be shared with the CUDA path
    d = init_net(RD.DeepConvDiscNet(40, 1), "cpu", "xavier")
    gt = [(lambda p, y, s, c: (p, y))(*inputs(B, N, 31 + i)) for i in range(iters)]
    nogt = [inputs(B, N, 131 + i)[0] for i in range(iters)]
    with torch.no_grad():
        g.eval()
        d_probe = d(F.log_softmax(g(nogt[2])[0], dim=1))
    TH, gap = pick_threshold(d_probe)
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(d.parameters(), lr=1e-5, betas=(0.9, 0.999))
    tp, ty, _, _ = inputs(8, N, 999)
    a = argparse.Namespace(device="cpu", total_iterations=iters, iter_save_epoch=1000, iter_test_epoch=1000,
                           tensorboard=False, exp_dir=tempfile.mkdtemp(), batch_size=B, input_pts=N,
                           lambda_cls=1.0, lambda_adv=1e-3, lambda_semi=1.0, semi_start=1, semi_TH=TH)
    torch.manual_seed(4242)
    RT.run_training_semi(gt, nogt, enumerate(gt), enumerate(nogt), [(tp[:B], ty[:B])], g, d,
                         torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(),
                         torch.nn.CrossEntropyLoss(ignore_index=255), opt, optD, ImagePool(0), ImagePool(0),
                         logger, logger, None, a)
    G["cls_semi"] = dict(recipe=dict(B=B, N=N, iters=iters, seed=31, seed2=131, wseed=0, label_seed=4242,
                                     semi_start=1, semi_TH=TH, TH_gap=gap), g=state(g), d=state(d))
    print("cls_semi done: TH %.6f (gap %.2e)" % (TH, gap))

    # ---------------------------------------------------------------- run_training_seg_semi
    B, N, iters = 2, 256, 3
    torch.manual_seed(0)
    g = init_net(PointNetSeg(50), "cpu", "xavier")
    d = init_net(RD.PointwiseDiscNet(N, 50), "cpu", "xavier")
    gt = [(lambda p, y, s, c: (p, c, s))(*inputs(B, N, 1234 + i)) for i in range(iters)]
    nogt = [(lambda p, y, s, c: (p, c))(*inputs(B, N, 4321 + i)) for i in range(iters)]
    import copy

    def run(g_, d_, n_iter, TH_):
        opt = torch.optim.Adam(g_.parameters(), lr=1e-4, betas=(0.9, 0.999))
        optD = torch.optim.Adam(d_.parameters(), lr=1e-5, betas=(0.9, 0.999))
        a = argparse.Namespace(device="cpu", total_iterations=n_iter, iter_save_epoch=1000, iter_test_epoch=1000,
                               tensorboard=False, exp_dir=tempfile.mkdtemp(), batch_size=B, input_pts=N,
                               lambda_seg=1.0, lambda_adv=1e-3, lambda_semi=1.0, semi_start=1, semi_TH=TH_)
        torch.manual_seed(4242)
        RT.run_training_seg_semi(gt, nogt, enumerate(gt), enumerate(nogt), seg_test(N), list(range(16)), g_, d_,
                                 torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss(),
                                 torch.nn.CrossEntropyLoss(ignore_index=255), opt, optD, ImagePool(0),
                                 ImagePool(0), logger, logger, None, a)

    # the threshold is placed in a gap of the discriminator outputs the THIRD iteration will see:
    # a dry run of the first two iterations (which do not depend on it) provides that state
    g0, d0 = copy.deepcopy(g), copy.deepcopy(d)
    run(g0, d0, 2, 0.0)
    with torch.no_grad():
        g0.train()
        d_probe = d0(F.log_softmax(g0(*nogt[2])[0], dim=1))
    TH, gap = pick_threshold(d_probe)
    run(g, d, iters, TH)
    G["seg_semi"] = dict(recipe=dict(B=B, N=N, iters=iters, seed=1234, seed2=4321, wseed=0, label_seed=4242,
                                     semi_start=1, semi_TH=TH, TH_gap=gap), g=state(g), d=state(d))
    print("seg_semi done: TH %.6f (gap %.2e)" % (TH, gap))
    torch.save(G, args.out)
    print("wrote", args.out, os.path.getsize(args.out) / 1e6, "MB")


if __name__ == "__main__":
    main()

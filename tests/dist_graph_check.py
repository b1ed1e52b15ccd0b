"""Data-parallel parity of the benched object (run under torch.distributed.run, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 \
        --master-port 29731 tests/dist_graph_check.py

Every rank builds the same G / D, takes its shard of a global batch and runs
``GraphedAdversarialSegStep(fused=True)`` with ``parallel.DistributedOptimizer`` -- the NCCL
all-reduce of the gradient slabs is captured inside the CUDA graph, the generator's exchange
overlapping the discriminator phase.  Checks:

* replicas hold bit-identical parameters after every step;
* they equal the parameters of ONE process stepping on the global batch with the same labels
  (eager ``adversarial_seg_step_fused``, plain Adam), to fp32 rounding.

Smoothed GAN labels are drawn once globally (seeded generator) and sharded like the batch, because
``make_D_label`` draws from each rank's own CPU generator (utils/utils.py:28).
Prints one JSON line and ``DIST_GRAPH_CHECK OK``.
"""
import argparse
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from adversarial_learning_on_pointclouds_b200 import models as M, Precision
    from adversarial_learning_on_pointclouds_b200.parallel import DistributedOptimizer, shard_batch
    from adversarial_learning_on_pointclouds_b200.trainer import (GraphedAdversarialSegStep,
                                                                  adversarial_seg_step_fused)
    from adversarial_learning_on_pointclouds_b200.utils import init_net
    from helpers import inputs, randomize_biases, rel_err

    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--clouds-per-rank", type=int, default=4)
    ap.add_argument("--points", type=int, default=512)
    ap.add_argument("--iters", type=int, default=4)
    a = ap.parse_args()
    world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Bglob, N = a.clouds_per_rank * world, a.points

    torch.manual_seed(1)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
    randomize_biases([g, d], 3)
    g.precision = d.precision = Precision(a.precision)
    g_ref, d_ref = copy.deepcopy(g), copy.deepcopy(d)
    g.to(dev); d.to(dev)
    mk = lambda m, lr: torch.optim.Adam(m.parameters(), lr=lr, betas=(0.9, 0.999), fused=True, capturable=True)
    opt, optD = DistributedOptimizer(mk(g, 1e-4)), DistributedOptimizer(mk(d, 1e-5))
    targs = argparse.Namespace(device=dev, lambda_seg=1.0, lambda_adv=1e-3)
    gan, ce = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()

    batches, labels = [], []
    lg = torch.Generator().manual_seed(99)
    for it in range(a.iters):
        pts, _, seg, cls = inputs(Bglob, N, 100 + it)
        pts2, _, _, cls2 = inputs(Bglob, N, 500 + it)
        batches.append(((pts, cls, seg), (pts2, cls2)))
        labels.append((torch.empty(Bglob, N).uniform_(0.7, 1.05, generator=lg),
                       torch.empty(Bglob, N).uniform_(0.0, 0.305, generator=lg)))
    mine = [(shard_batch(bg, rank, world), shard_batch(bn, rank, world)) for bg, bn in batches]
    my_labels = [shard_batch(l, rank, world) for l in labels]
    drawn = [0]

    def label_draw(real, fake):
        i = min(drawn[0], a.iters - 1)
        real.copy_(my_labels[i][0]); fake.copy_(my_labels[i][1])
        drawn[0] += 1

    to_dev = lambda part: tuple(t.to(dev) for t in part)
    gstep = GraphedAdversarialSegStep(g, d, gan, ce, opt, optD, targs, to_dev(mine[0][0]), to_dev(mine[0][1]),
                                      warmup=2, fused=True, label_draw=label_draw)
    worst_replica = 0.0
    losses = []
    for it in range(a.iters):
        l = gstep(to_dev(mine[it][0]), to_dev(mine[it][1]))
        losses.append(l.clone())
        flat = torch.cat([p.detach().reshape(-1) for p in list(g.parameters()) + list(d.parameters())])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        for other in gathered[1:]:
            worst_replica = max(worst_replica, (other - gathered[0]).abs().max().item())
    torch.cuda.synchronize()
    assert worst_replica == 0.0, "replicas diverged: %g" % worst_replica
    modes = (opt.last_mode, optD.last_mode)

    result = {"world": world, "precision": a.precision, "iters": a.iters, "replica_max_abs_diff": worst_replica,
              "exchange_mode": modes, "launches_per_step": gstep.launches_per_step}
    if rank == 0:
        # one process, global batch, same labels, eager loop body, plain Adam
        g_ref.to(dev); d_ref.to(dev)
        ropt, roptD = mk(g_ref, 1e-4), mk(d_ref, 1e-5)
        worst_loss = 0.0
        for it in range(a.iters):
            lab = tuple(t.to(dev) for t in labels[it])
            fn = lambda d_out, value, rnd, lab=lab: (torch.full_like(d_out, float(value)) if not rnd
                                                     else (lab[0] if value == 1 else lab[1]))
            adversarial_seg_step_fused(g_ref, d_ref, gan, ce, ropt, roptD, to_dev(batches[it][0]),
                                       to_dev(batches[it][1]), targs, label_fn=fn)
        errs = {}
        for (k, p), (_, q) in zip(list(g.named_parameters()) + list(d.named_parameters()),
                                  list(g_ref.named_parameters()) + list(d_ref.named_parameters())):
            errs[k] = rel_err(p, q)
        result["worst_param_rel_err_vs_global_batch"] = max(errs.values())
        result["rank0_losses_last"] = [float(x) for x in losses[-1].cpu()]
        print(json.dumps(result), flush=True)
        # fp32 rounding, plus -- seen at 8 ranks x 2 clouds -- one decision that lands on the other side
        # between the shard-sized and the global-batch launch (different kernels below 1024 rows)
        assert max(errs.values()) < (1e-4 if a.precision == "fp32" else 1e-3), errs
        print("DIST_GRAPH_CHECK OK", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)            # see bench.py: tearing NCCL down under a live captured graph can block


if __name__ == "__main__":
    main()

"""CPU, world_size 2 over gloo: the data-parallel exchange (one flat gradient
all-reduce per optimizer step).  The kernels need a GPU, so the
model here is the CPU oracle; what is under test is the host logic of
parallel.DistributedOptimizer / shard_batch: averaged shard gradients equal the
global-batch gradient, None gradients count as zero, parameters stay identical
on both ranks after the step."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _loss(params, pts, cls, seg):
    from oracle import pointnet_oracle as PO
    pred, _ = PO.pointnet_seg_forward(params, pts, cls)
    return F.cross_entropy(pred, seg)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from adversarial_learning_on_pointclouds_b200.parallel import DistributedOptimizer, shard_batch
    from oracle import pointnet_oracle as PO, steps
    sd = PO.random_state_dict(PO.pointnet_seg_shapes(50), seed=5)
    params = steps.leaf_params(sd)
    unused = torch.zeros(7, requires_grad=True)                  # never receives a gradient
    opt = DistributedOptimizer(torch.optim.SGD(list(params.values()) + [unused], lr=0.1))
    pts, _, seg, cls = PO.synthetic_inputs(4, 64, 11)            # global batch of 4 clouds
    my = shard_batch((pts, cls, seg), rank, world)
    opt.zero_grad()
    _loss(params, *my).backward()
    opt.step()
    # reference: one process, global batch
    ref = steps.leaf_params(sd)
    _loss(ref, pts, cls, seg).backward()
    worst = 0.0
    for k in params:
        want = sd[k] - 0.1 * ref[k].grad
        worst = max(worst, ((params[k].detach() - want).norm() / want.norm().clamp_min(1e-12)).item())
    assert worst < 1e-5, worst
    assert opt.last_mode == "packed"                             # torch autograd: one storage per gradient
    assert unused.grad is None                                   # as in a single process: Adam / SGD skip it
    # ---- second iteration with zero_grad(set_to_none=False): gradients accumulate into the
    # same tensors, which must not alias the exchange buffer
    before = {k: v.detach().clone() for k, v in params.items()}
    opt.zero_grad(set_to_none=False)
    _loss(params, *my).backward()
    opt.reduce_gradients_async()                                 # started early, as the step functions do
    opt.reduce_gradients_async()                                 # idempotent
    opt.step()
    ref2 = steps.leaf_params(before)
    _loss(ref2, pts, cls, seg).backward()
    for k in params:
        want = before[k] - 0.1 * ref2[k].grad
        worst = max(worst, ((params[k].detach() - want).norm() / want.norm().clamp_min(1e-12)).item())
    assert worst < 1e-5, worst
    # ---- in-place exchange: gradients that are views of one zero-filled slab (what the libpcadv
    # backward hands autograd) are reduced where they lie
    ws = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7))]
    opt2 = DistributedOptimizer(torch.optim.SGD(ws, lr=1.0))
    slab = torch.zeros(40)
    ws[0].grad = slab[0:15].view(5, 3)
    ws[1].grad = slab[16:23]
    slab[0:15] = float(rank + 1)
    slab[16:23] = float(10 * (rank + 1))
    opt2.step()
    assert opt2.last_mode == "in_place"
    assert ws[0].grad.untyped_storage().data_ptr() == slab.untyped_storage().data_ptr()
    assert torch.allclose(ws[0].detach(), torch.full((5, 3), -1.5)) and torch.allclose(ws[1].detach(), torch.full((7,), -15.0))
    flat = torch.cat([p.detach().reshape(-1) for p in params.values()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert torch.equal(gathered[0], gathered[1])                 # replicas stay identical
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("%.3e" % worst)
    dist.destroy_process_group()


def test_distributed_optimizer_matches_global_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_shard_batch_shapes():
    from adversarial_learning_on_pointclouds_b200.parallel import shard_batch
    a, b = torch.arange(8).view(8, 1), torch.arange(16).view(8, 2)
    s0, s1 = shard_batch((a, b), 0, 2), shard_batch((a, b), 1, 2)
    assert torch.equal(torch.cat([s0[0], s1[0]]), a) and torch.equal(torch.cat([s0[1], s1[1]]), b)
    try:
        shard_batch((a,), 0, 3)
    except ValueError:
        return
    raise AssertionError("uneven shard must raise")

"""Data parallelism for the adversarial step: one process per GPU, parameters
replicated, the batch sharded over ranks, and ONE exchange per optimizer step --
a sum all-reduce of a flat fp32 gradient buffer (NCCL over NVLink / NVSwitch),
divided by the world size.

The reference has no distributed code.  The trainer only ever calls
``optimizer.zero_grad()`` and ``optimizer.step()`` (utils/trainer.py:881-882,
:965-966), so the exchange hooks in there and the loop runs unchanged.
DistributedDataParallel is not used: the loop toggles ``requires_grad`` on D
every iteration (:885-886, :932-933), runs D three times forward and twice
backward per step, and BaseDiscNet.conv4 never receives a gradient -- each of
which trips DDP's reducer bookkeeping.  A ``None`` gradient counts as zero.
"""
import torch
import torch.distributed as dist


class DistributedOptimizer:
    """Wraps a torch optimizer; ``step()`` first averages gradients over ranks."""

    def __init__(self, optimizer, process_group=None):
        self.optimizer = optimizer
        self.group = process_group
        self.params = [p for g in optimizer.param_groups for p in g["params"]]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=torch.float32, device=ref.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, sd):
        self.optimizer.load_state_dict(sd)

    def zero_grad(self, set_to_none=True):
        self.optimizer.zero_grad(set_to_none=set_to_none)

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def reduce_gradients(self):
        world = self.world_size()
        if world == 1:
            return
        have = [p.grad is not None for p in self.params]
        self.flat.zero_()
        src = [p.grad for p, h in zip(self.params, have) if h]
        dst = [v for v, h in zip(self.views, have) if h]
        if src:
            torch._foreach_copy_(dst, src)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / world)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def step(self, closure=None):
        self.reduce_gradients()
        return self.optimizer.step(closure) if closure is not None else self.optimizer.step()


def shard_batch(tensors, rank, world):
    """This rank's equal slice of a global batch along dim 0."""
    out = []
    for t in tensors:
        b = t.shape[0]
        if b % world:
            raise ValueError("global batch %d is not divisible by world size %d" % (b, world))
        per = b // world
        out.append(t[rank * per:(rank + 1) * per])
    return tuple(out)

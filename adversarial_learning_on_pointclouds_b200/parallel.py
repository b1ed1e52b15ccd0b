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
    """Wraps a torch optimizer; ``step()`` first averages gradients over ranks.

    Where the gradients live decides how they travel:

    * **in place** -- the backward of the libpcadv Functions carves every dW / dbias accumulator of a
      pass out of ONE zero-filled fp32 slab (``models/_chain.ZeroPool``) and autograd adopts those
      views as ``p.grad``.  When the gradients of this optimizer sit in a handful of such slabs, the
      slabs themselves are all-reduced: no zero-fill, no gather copy, no scatter back.
    * **packed** -- otherwise (gradients produced by torch ops, many storages) they are copied into
      one flat buffer, reduced, and copied back into the ``p.grad`` tensors (which are never made to
      alias the buffer, so ``zero_grad(set_to_none=False)`` keeps working).

    ``reduce_gradients_async()`` launches the exchange and returns at once; the step functions of
    ``trainer.py`` call it for the generator right after its backward, so that all-reduce runs under
    the discriminator phase (NCCL's own stream; inside a CUDA-graph capture it becomes a parallel
    branch of the graph).  ``step()`` starts the exchange if nobody did, waits for it, and steps.
    """

    MAX_SLABS = 8

    def __init__(self, optimizer, process_group=None):
        self.optimizer = optimizer
        self.group = process_group
        self.params = [p for g in optimizer.param_groups for p in g["params"]]
        self.flat = None                 # packed path only, allocated on first use
        self._pending = None
        self.last_mode = None            # "in_place" | "packed": what the last exchange did (tests, bench)

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    @property
    def state(self):
        return self.optimizer.state

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, sd):
        self.optimizer.load_state_dict(sd)

    def zero_grad(self, set_to_none=True):
        self.optimizer.zero_grad(set_to_none=set_to_none)

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _avg_op(self):
        """(reduce op, factor still to be applied): NCCL averages inside the collective."""
        if dist.get_backend(self.group) == "nccl":
            return dist.ReduceOp.AVG, None
        return dist.ReduceOp.SUM, 1.0 / self.world_size()

    def _slabs(self, grads):
        """The storages behind ``grads`` as flat fp32 tensors, or None when the in-place exchange does
        not apply (non-fp32, too many storages, or storages much larger than the gradients)."""
        seen = {}
        for g in grads:
            if g.dtype != torch.float32 or not g.is_contiguous():
                return None
            st = g.untyped_storage()
            seen.setdefault(st.data_ptr(), (st, g))
            if len(seen) > self.MAX_SLABS:
                return None
        total = sum(g.numel() for g in grads)
        held = sum(st.nbytes() // 4 for st, _ in seen.values())
        if held > 2 * total + (1 << 16):
            return None
        out = []
        for st, g in seen.values():
            out.append(torch.empty(0, dtype=torch.float32, device=g.device).set_(st, 0, (st.nbytes() // 4,)))
        return out

    def reduce_gradients_async(self):
        """Launch the gradient exchange of this optimizer (idempotent until ``step()``)."""
        if self._pending is not None or self.world_size() == 1:
            return
        have = [p for p in self.params if p.grad is not None]
        grads = [p.grad for p in have]
        op, post = self._avg_op()
        slabs = self._slabs(grads) if grads else []
        if slabs is not None:
            self.last_mode = "in_place"
            works = [dist.all_reduce(t, op=op, group=self.group, async_op=True) for t in slabs]
            self._pending = (works, slabs, post, None)
            return
        self.last_mode = "packed"
        n = sum(g.numel() for g in grads)
        if self.flat is None or self.flat.numel() < n:
            self.flat = torch.empty(sum(p.numel() for p in self.params), dtype=torch.float32,
                                    device=grads[0].device)
        views, off = [], 0
        for g in grads:
            views.append(self.flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(views, grads)
        buf = self.flat[:n]
        work = dist.all_reduce(buf, op=op, group=self.group, async_op=True)
        self._pending = ([work], [buf], post, (grads, views))

    def reduce_gradients(self):
        """Exchange now and wait (what ``step()`` does when nothing was started)."""
        self.reduce_gradients_async()
        if self._pending is None:
            return
        works, bufs, post, unpack = self._pending
        self._pending = None
        for w in works:
            w.wait()
        if post is not None:
            for t in bufs:
                t.mul_(post)
        if unpack is not None:
            grads, views = unpack
            torch._foreach_copy_(grads, views)

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

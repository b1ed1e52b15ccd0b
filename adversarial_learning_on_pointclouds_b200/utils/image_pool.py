"""History pool of generated maps behind the reference's ``ImagePool`` interface
(utils/image_pool.py:10-55), kept on the device.

The reference keeps a Python list of 1-sample tensors and, for every sample of a batch, appends,
clones or swaps one of them, then concatenates the batch again: 2-3 tiny kernels and a list
operation per cloud.  Here the pool is ONE preallocated device tensor ``[pool_size, *sample]``.
The per-sample decisions are still taken on the host with Python's ``random`` module, in the
reference's order (one ``uniform(0, 1)`` per sample once the pool is full, one ``randint`` per
swap), so a seeded run makes the same choices -- but they only produce index lists, and the data
moves in at most three batched index kernels per query.
"""
import random

import torch


class ImagePool:
    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.num_imgs = 0
        self.store = None                      # [pool_size, *sample shape], allocated on first use

    @property
    def images(self):
        """The stored samples as the reference's list of 1-sample tensors (read-only view)."""
        if self.store is None:
            return []
        return [self.store[i:i + 1] for i in range(self.num_imgs)]

    def _plan(self, batch):
        """Replay the reference's per-sample loop on indices only.

        Returns (src, fill): src[i] = ("in", k) if output i is input sample k, ("pool", j) if it is
        what slot j held BEFORE this query; fill[j] = k for every slot whose final content is input
        sample k.  A slot written earlier in the same batch can be handed out again later in it
        (the reference's list is updated in place), hence the ``now`` table."""
        now = {}                               # slot -> input index currently stored there
        src = []
        for i in range(batch):
            if self.num_imgs < self.pool_size:
                now[self.num_imgs] = i
                self.num_imgs += 1
                src.append(("in", i))
            elif random.uniform(0, 1) > 0.5:
                j = random.randint(0, self.pool_size - 1)
                src.append(("in", now[j]) if j in now else ("pool", j))
                now[j] = i
            else:
                src.append(("in", i))
        return src, now

    def query(self, images):
        """pool_size 0: the input itself.  Otherwise a new leaf tensor (``requires_grad=True``, as
        utils/image_pool.py:53-55) holding, per sample, either the input or a stored older one."""
        if self.pool_size == 0:
            return images
        data = images.detach()
        if self.store is None:
            self.store = torch.empty((self.pool_size,) + tuple(data.shape[1:]), dtype=data.dtype,
                                     device=data.device)
        src, fill = self._plan(data.shape[0])
        dev = data.device
        out = data.clone()
        moved = [(i, k) for i, (kind, k) in enumerate(src) if kind == "in" and k != i]
        if moved:                              # a sample stored and handed out again within the batch
            dst, frm = zip(*moved)
            out[torch.tensor(dst, device=dev)] = data[torch.tensor(frm, device=dev)]
        old = [(i, j) for i, (kind, j) in enumerate(src) if kind == "pool"]
        if old:                                # read the old slots before they are overwritten
            dst, slots = zip(*old)
            out[torch.tensor(dst, device=dev)] = self.store[torch.tensor(slots, device=dev)]
        if fill:
            slots, frm = zip(*fill.items())
            self.store[torch.tensor(slots, device=dev)] = data[torch.tensor(frm, device=dev)]
        return out.requires_grad_(True)

"""Per-sample restatement of the reference's history pool (utils/image_pool.py:26-55) on sample
IDs instead of tensors.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``); pinned by
``tests/golden/image_pool_golden.json`` (outputs of the reference's own ``ImagePool``)."""
import random


class IdPool:
    """Tracks which sample ID each slot holds; ``query(ids)`` returns the IDs the reference would
    return for a batch of samples carrying those IDs, consuming ``random`` exactly as it does."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.slots = []

    def query(self, ids):
        if self.pool_size == 0:                                   # :35-36
            return list(ids)
        out = []
        for s in ids:                                             # :38
            if len(self.slots) < self.pool_size:                  # :40-43
                self.slots.append(s)
                out.append(s)
            elif random.uniform(0, 1) > 0.5:                      # :45-46
                j = random.randint(0, self.pool_size - 1)         # :47
                out.append(self.slots[j])                         # :48, :50
                self.slots[j] = s                                 # :49
            else:                                                 # :51-52
                out.append(s)
        return out

#!/usr/bin/env python3
"""Generate tests/golden/image_pool_golden.json with the UNMODIFIED reference ``ImagePool``
(utils/image_pool.py, imported read-only from /root/reference).  Build container only.

Each sample is a tensor filled with its own serial number, so the returned batch spells out which
sample the pool handed back."""
import json
import os
import random
import sys

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    from utils.image_pool import ImagePool
    cases = []
    for pool_size, batch, rounds, seed in [(0, 4, 3, 1), (5, 4, 12, 1234), (3, 8, 10, 7), (16, 3, 30, 99),
                                           (1, 6, 6, 5)]:
        random.seed(seed)
        pool = ImagePool(pool_size)
        serial, outs = 0, []
        for _ in range(rounds):
            ids = list(range(serial, serial + batch))
            serial += batch
            x = torch.tensor(ids, dtype=torch.float32).view(batch, 1, 1).expand(batch, 2, 3).contiguous()
            y = pool.query(x)
            assert y.shape == x.shape
            outs.append([int(v) for v in y[:, 0, 0].tolist()])
        cases.append({"pool_size": pool_size, "batch": batch, "rounds": rounds, "seed": seed, "out": outs})
    with open(os.path.join(HERE, "image_pool_golden.json"), "w") as f:
        json.dump(cases, f)
    print("wrote image_pool_golden.json:", [(c["pool_size"], c["batch"], c["rounds"]) for c in cases])


if __name__ == "__main__":
    main()

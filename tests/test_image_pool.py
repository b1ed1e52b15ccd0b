"""History pool: the ID-level oracle and the package's batched device pool
against outputs of the reference's own ``ImagePool`` (tests/golden/image_pool_golden.json, made by
tests/golden/make_image_pool_golden.py).  Exact: the pool only moves data."""
import json
import os
import random

import pytest
import torch

from adversarial_learning_on_pointclouds_b200.utils import ImagePool
from oracle.image_pool_oracle import IdPool

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "image_pool_golden.json")) as f:
    CASES = json.load(f)
IDS = ["pool%d_b%d" % (c["pool_size"], c["batch"]) for c in CASES]


def _batches(case):
    serial = 0
    for _ in range(case["rounds"]):
        yield list(range(serial, serial + case["batch"]))
        serial += case["batch"]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_pool_matches_reference(case):
    random.seed(case["seed"])
    pool = IdPool(case["pool_size"])
    assert [pool.query(ids) for ids in _batches(case)] == case["out"]


def _run_pool(case, device):
    random.seed(case["seed"])
    pool = ImagePool(case["pool_size"])
    outs = []
    for ids in _batches(case):
        x = torch.tensor(ids, dtype=torch.float32, device=device).view(-1, 1, 1).expand(len(ids), 2, 3)
        x = x.contiguous().requires_grad_(True) * 1.0          # a non-leaf, like a generator output
        y = pool.query(x)
        assert y.shape == x.shape and y.device == x.device
        if case["pool_size"] > 0:
            assert y.is_leaf and y.requires_grad                # utils/image_pool.py:54
            assert (y == y[:, :1, :1]).all()                    # whole samples move together
        outs.append([int(v) for v in y[:, 0, 0].tolist()])
    stored = sorted(int(t[0, 0, 0]) for t in pool.images)
    return outs, stored


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_batched_pool_matches_reference_host_tensors(case):
    outs, stored = _run_pool(case, "cpu")
    assert outs == case["out"]
    random.seed(case["seed"])
    ref = IdPool(case["pool_size"])
    for ids in _batches(case):
        ref.query(ids)
    assert stored == sorted(ref.slots)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_batched_pool_matches_reference_on_device(case):
    outs, _ = _run_pool(case, "cuda")
    assert outs == case["out"]

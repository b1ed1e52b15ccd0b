"""CPU oracle for the PointNet + discriminator adversarial train step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
directory: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the timed CPU baseline, never as the thing shipped.

The oracle is a functional restatement (state-dict in, tensors out) of the
reference's ``models/pointnet.py`` and ``models/discriminator.py`` forward
passes and of the loop bodies in ``utils/trainer.py``, executed by the same
arithmetic library the reference uses on CPU (ATen fp32).  Every function cites
the reference lines it follows.

Parity pinning: the reference ships no golden vectors or tests,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference`` unmodified) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks the oracle against those fixtures.
"""
from . import pointnet_oracle, discriminator_oracle, steps  # noqa: F401

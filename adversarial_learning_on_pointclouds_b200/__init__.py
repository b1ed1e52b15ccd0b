"""B200-native PointNet + discriminator adversarial train step.

Drop-in for the hot path of YiruS/Adversarial_Learning_on_PointClouds: the
``models/pointnet.py`` and ``models/discriminator.py`` module API, backed by
hand-written sm_100a kernels in ``libpcadv.so`` (C ABI in ``include/pcadv.h``).
CUDA only -- there is no CPU or PyTorch-eager fallback.
"""
from . import _lib, ops                                   # noqa: F401
from .ops import Precision, set_default_precision, default_precision   # noqa: F401
from . import models                                      # noqa: F401

__version__ = "0.1.0"

"""CPU restatement of ``pcadv_jitter`` (TEST INFRASTRUCTURE, see ``oracle/__init__.py``).

The reference jitters on the host with numpy's generator -- ``np.clip(sigma * np.random.randn(N, C),
-clip, clip) + data`` (dataset/modelNetData.py:80-91) -- whose stream cannot be reproduced on the
device.  The device kernel therefore defines its own counter-based stream (Philox-4x32-10, Box-Muller
on 24-bit uniforms); this file restates that stream in numpy so the kernel can be checked element by
element, and ``reference_jitter`` restates the reference's formula for the distribution checks.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, seed):
    """counter: uint64 array; seed: int.  Returns uint32 [len(counter), 4]."""
    counter = np.asarray(counter, dtype=np.uint64)
    c = [(counter & _MASK).astype(np.uint32), (counter >> np.uint64(32)).astype(np.uint32),
         np.zeros_like(counter, dtype=np.uint32), np.zeros_like(counter, dtype=np.uint32)]
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[0].astype(np.uint64)
            p1 = _M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return np.stack(c, axis=1)


def _u01(x):
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def normals(count, seed, offset=0):
    """The first ``count`` standard normals of the device stream (float32)."""
    quads = (count + 3) // 4
    r = philox4x32_10(np.arange(quads, dtype=np.uint64) + np.uint64(offset), seed)
    z = np.empty((quads, 4), dtype=np.float32)
    for h in range(2):
        rad = np.sqrt(np.float32(-2.0) * np.log(_u01(r[:, 2 * h]))).astype(np.float32)
        ang = (np.float32(6.283185307179586) * _u01(r[:, 2 * h + 1])).astype(np.float32)
        z[:, 2 * h] = rad * np.cos(ang)
        z[:, 2 * h + 1] = rad * np.sin(ang)
    return z.reshape(-1)[:count]


def jitter(data, sigma=0.01, clip=0.05, seed=0, offset=0):
    """``pcadv_jitter`` on a float32 array of any shape."""
    flat = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    z = normals(flat.size, seed, offset)
    out = flat + np.clip(np.float32(sigma) * z, np.float32(-clip), np.float32(clip))
    return out.reshape(np.shape(data)).astype(np.float32)


def reference_jitter(data, sigma=0.01, clip=0.05, rng=None):
    """jitter_point_cloud of dataset/modelNetData.py:80-91 (numpy's generator on the host)."""
    rng = rng or np.random
    N, C = data.shape
    assert clip > 0
    jittered = np.clip(sigma * rng.randn(N, C), -1 * clip, clip)
    jittered += data
    return jittered

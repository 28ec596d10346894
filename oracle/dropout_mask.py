"""ORACLE -- test infrastructure, not product code.

numpy restatement of the kernels' stateless dropout mask (csrc/common.cuh: mix32 / drop_rowseed / drop_pair /
make_drop). The reference uses torch's nn.Dropout, whose random stream no other implementation can reproduce, so
train()-mode parity is defined as: "the kernels compute exactly what the reference computes when the reference's
nn.Dropout masks are REPLACED by these masks" -- the oracle (fcmf_oracle.py) takes the masks from here and the
kernels regenerate the same bits on the device. ``tests/test_cpu_host.py`` pins this file against the C++ functions
themselves (compiled for the host from common.cuh's definitions)."""
from __future__ import annotations

import numpy as np

GOLDEN64 = 0x9E3779B97F4A7C15
M64 = (1 << 64) - 1


def mix32(x):
    """'lowbias32' integer finaliser on uint32 arrays (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
    return x


def splitmix64(z: int) -> int:
    z = (z + GOLDEN64) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def rowseed(seed: int, rows) -> np.ndarray:
    rows = np.asarray(rows, dtype=np.uint64)
    seed = splitmix64(seed & M64)
    lo, hi = np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32)
    a = mix32((rows & np.uint64(0xFFFFFFFF)).astype(np.uint32) ^ lo)
    with np.errstate(over="ignore"):
        b = (rows >> np.uint64(32)).astype(np.uint32) * np.uint32(0x9E3779B9) + hi
    return mix32(a ^ b)


def threshold(p: float):
    """(thr16, inv_keep) exactly as make_drop computes them in float32."""
    p32 = np.float32(max(p, 0.0))
    thr = int(np.uint32(p32 * np.float32(65536.0) + np.float32(0.5)))
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(thr) * np.float32(1.0 / 65536.0))
    return thr, float(inv)


def keep_mask(seed: int, rows, ncols: int, p: float) -> np.ndarray:
    """bool [len(rows), ncols]: True where the element survives dropout."""
    thr, _ = threshold(p)
    rs = rowseed(seed, rows).reshape(-1, 1)
    cols = np.arange(ncols, dtype=np.uint32).reshape(1, -1)
    with np.errstate(over="ignore"):
        h = mix32(rs + (cols >> np.uint32(1)))
    hw = np.where((cols & np.uint32(1)).astype(bool), h >> np.uint32(16), h & np.uint32(0xFFFF))
    return hw >= np.uint32(thr)


def site_seed(step_seed: int, site: int) -> int:
    return (step_seed + site * GOLDEN64) & M64


def scaled_mask(seed: int, rows, ncols: int, p: float):
    """float32 numpy [len(rows), ncols]: keep / (1 - p_quantised)  (what nn.Dropout multiplies by)."""
    _, inv = threshold(p)
    return keep_mask(seed, rows, ncols, p).astype(np.float32) * np.float32(inv)

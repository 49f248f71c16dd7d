"""Exact IEEE-754 binary32 helpers for the oracle (test infrastructure only).

numpy never contracts a*b+c into an FMA, so plain float32 numpy expressions are
the "separately rounded" semantics (CUDA: __fmul_rn / __fadd_rn).  Where the
upstream CUDA code uses an explicit or compiler-contracted FMA we need a
correctly rounded single-precision fma on the CPU; `fma32` provides it.
"""
import numpy as np

F32 = np.float32


def f32(x):
    return np.asarray(x, dtype=np.float32)


def fma32(a, b, c):
    """Correctly rounded float32 fma(a, b, c) for float32 array inputs (vectorised, exact).

    p = a*b is exact in float64 (24+24 <= 53 bits).  s = fl64(p + c) may round; TwoSum recovers the
    exact residual err = (p + c) - s.  Casting s to float32 is a correct single rounding unless s
    sits exactly on a float32 tie (low 29 mantissa bits == 0x10000000) while err != 0: the true
    value is then strictly on one side of the tie and the result is the neighbour on that side.
    """
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    c = np.asarray(c, dtype=np.float32)
    a, b, c = np.broadcast_arrays(a, b, c)
    with np.errstate(all="ignore"):
        p = a.astype(np.float64) * b.astype(np.float64)
        c64 = c.astype(np.float64)
        s = p + c64
        bb = s - p
        err = (p - (s - bb)) + (c64 - bb)
        out = s.astype(np.float32)
        bits = np.ascontiguousarray(s).view(np.uint64)
        tie = ((bits & np.uint64(0x1FFFFFFF)) == np.uint64(0x10000000)) & np.isfinite(s) & (err != 0)
        if np.any(tie):
            near = out.astype(np.float64)
            up = np.where(near > s, out, np.nextafter(out, np.float32(np.inf)))
            dn = np.where(near > s, np.nextafter(out, np.float32(-np.inf)), out)
            out = np.where(tie, np.where(err > 0, up, dn), out).astype(np.float32)
    return out


def lowbias32(x):
    """Counter-based 32-bit mixer (C. Wellons' lowbias32) on uint32 arrays."""
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
    return x


def jitter_u01(seed: int, idx):
    """u in [0,1): the marcher's jitter stream.  idx = flat (ray*S+step) or (nugget*S+step).

    Definition shared with csrc/common.cuh::pag_jitter: lowbias32(idx + seed*0x9E3779B9) >> 8, * 2^-24.
    """
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (idx + np.uint64(seed) * np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)
    h = lowbias32(x.astype(np.uint32))
    return (h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)

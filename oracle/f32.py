"""Exact IEEE-754 binary32 helpers for the oracle (test infrastructure only).

numpy never contracts a*b+c into an FMA, so plain float32 numpy expressions are
the "separately rounded" semantics (CUDA: __fmul_rn / __fadd_rn).  Where the
upstream CUDA code uses an explicit or compiler-contracted FMA we need a
correctly rounded single-precision fma on the CPU; `fma32` provides it.
"""
from fractions import Fraction

import numpy as np

F32 = np.float32


def f32(x):
    return np.asarray(x, dtype=np.float32)


def fma32(a, b, c):
    """Correctly rounded float32 fma(a, b, c) for float32 array inputs.

    a*b is exact in float64 (24+24 <= 53 bits).  The float64 add rounds once
    and the cast to float32 rounds again; the two roundings can only disagree
    with a true fma when the float64 sum lies exactly on a float32 rounding
    boundary (low 29 mantissa bits == 0x10000000).  Those (astronomically rare)
    elements are recomputed with exact rational arithmetic.
    """
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    c = np.asarray(c, dtype=np.float32)
    a, b, c = np.broadcast_arrays(a, b, c)
    with np.errstate(all="ignore"):
        s = a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)
        out = s.astype(np.float32)
    bits = s.view(np.uint64) if s.flags["C_CONTIGUOUS"] else np.ascontiguousarray(s).view(np.uint64)
    tie = ((bits & np.uint64(0x1FFFFFFF)) == np.uint64(0x10000000)) & np.isfinite(s)
    if np.any(tie):
        out = out.copy()
        idx = np.nonzero(tie)
        for i in zip(*idx):
            exact = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
            out[i] = _round_fraction_to_f32(exact)
    return out


def _round_fraction_to_f32(q: Fraction) -> np.float32:
    # round-to-nearest-even of an exact rational to binary32 via two candidates
    lo = np.float32(float(q))  # float(q) is correctly rounded to double; cast may double-round
    cands = [lo, np.nextafter(lo, np.float32(np.inf)), np.nextafter(lo, np.float32(-np.inf))]
    best = None
    for cnd in cands:
        err = abs(Fraction(float(cnd)) - q)
        key = (err, int(np.float32(cnd).view(np.uint32)) & 1)  # ties -> even mantissa
        if best is None or key < best[0]:
            best = (key, cnd)
    return np.float32(best[1])


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

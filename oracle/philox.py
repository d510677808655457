"""Philox4x32-10 in numpy, bit-identical to ssd_b200/csrc/common.cuh::philox4x32_10 and the
keying of ssd_b200/csrc/sampling.cuh::philox_draw.  Replaces the reference's use of torch's
global CUDA Philox stream (layers/sampler.py:6,33; utils/verify.py:115,158-159)."""
from __future__ import annotations

import numpy as np

TAG_SAMPLE, TAG_ACCEPT, TAG_RECOVER = 1, 2, 3
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Counters are uint32 arrays (broadcastable); key is two python ints. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def draw(idx, row, call_id: int, tag: int, seed: int):
    """sampling.cuh::philox_draw — counter = (idx, row, call_lo, call_hi24 | tag<<24), key = seed."""
    c2 = call_id & 0xFFFFFFFF
    c3 = ((call_id >> 32) & 0x00FFFFFF) | (tag << 24)
    idx = np.asarray(idx, dtype=np.uint64)
    return philox4x32_10(idx, np.asarray(row, dtype=np.uint64), np.uint64(c2), np.uint64(c3), seed & 0xFFFFFFFF,
                         (seed >> 32) & 0xFFFFFFFF)


def unit_open0(x):
    """(0,1]: ((x>>8)+1) * 2^-24 (common.cuh::u32_to_unit_open0)."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)


def unit_half_open(x):
    """[0,1) like torch.rand: (x>>8) * 2^-24 (sampling.cuh::u32_to_unit_half_open)."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def exp1(x):
    """Exp(1) sample: -log(u), u in (0,1]."""
    return (-np.log(unit_open0(x))).astype(np.float32)

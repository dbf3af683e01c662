"""numpy Philox4x32-10 (Salmon, Moraes, Dror, Shaw -- SC'11 "Parallel random numbers: as easy
as 1, 2, 3") and the action sampler built on it.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Mirrors the device sampler in
csrc/common.cuh (fetch_action) bit for bit so that plans made with device-sampled actions
can be re-scored by the float64 oracle.  Known-answer vectors from the Random123
distribution (kat_vectors) are checked in tests/test_philox.py.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
    """counter [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [np.asarray(counter[..., i], dtype=np.uint32) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32)
    k1 = np.asarray(key[..., 1], dtype=np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[0].astype(np.uint64)
            p1 = _M1 * c[2].astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c[1] ^ k0
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c[3] ^ k1
            n3 = p0.astype(np.uint32)
            c = [n0, n1, n2, n3]
            k0 = k0 + _W0
            k1 = k1 + _W1
    return np.stack(c, axis=-1)


def sample_actions(K, H, da, seed, low, high, k_offset=0):
    """actions [K, H, da] float64 exactly as the device sampler produces them:
    element e = t*da + j of sequence k uses word e%4 of Philox(counter=(k_lo, k_hi, e//4, 0),
    key=(seed_lo, seed_hi)); u = (word >> 8) * 2^-24; a = float32(low + u * (high - low))."""
    low = np.broadcast_to(np.asarray(low, dtype=np.float64), (da,))
    high = np.broadcast_to(np.asarray(high, dtype=np.float64), (da,))
    E = H * da
    nblk = (E + 3) // 4
    k = (np.arange(K, dtype=np.uint64) + np.uint64(k_offset))[:, None]
    blk = np.arange(nblk, dtype=np.uint64)[None, :]
    ctr = np.zeros((K, nblk, 4), dtype=np.uint32)
    ctr[..., 0] = (k & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[..., 1] = (k >> np.uint64(32)).astype(np.uint32)
    ctr[..., 2] = blk.astype(np.uint32)
    key = np.zeros((K, nblk, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    words = philox4x32_10(ctr, key).reshape(K, nblk * 4)[:, :E]
    u = (words >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)
    u = u.reshape(K, H, da)
    a = low[None, None, :] + u * (high - low)[None, None, :]
    return a.astype(np.float32).astype(np.float64)


# Random123 known-answer tests for philox4x32_10: (counter, key, expected)
kat_vectors = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]

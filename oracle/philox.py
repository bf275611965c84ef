"""ORACLE — test infrastructure only.

numpy restatement of the engine's seeded noise generator (csrc/elementwise.cu `randn_kernel`): Philox4x32-10
(Salmon et al., SC'11; same round constants as cuRAND / libtorch's Philox) + Box-Muller in f32. The reference
itself draws UNSEEDED libtorch noise (src/model/stablediffusion/mod.rs:378-388), so there is nothing to match
bit-for-bit there; this file only pins our own generator. The integer stream is bit-exact; the Box-Muller
floats differ from the GPU by libm ulps (tests use a small tolerance).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [n,4] uint32, key [2] uint32 -> [n,4] uint32."""
    c = ctr.astype(np.uint32).copy()
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    for _ in range(10):
        p0 = M0 * c[:, 0].astype(np.uint64)
        p1 = M1 * c[:, 2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        n0 = hi1 ^ c[:, 1] ^ k0
        n2 = hi0 ^ c[:, 3] ^ k1
        c = np.stack([n0, lo1, n2, lo0], axis=1)
        with np.errstate(over="ignore"):
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c


def philox_words(n: int, seed: int, subseq: int = 0) -> np.ndarray:
    nblk = (n + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64)
    ctr = np.stack([(blk & np.uint64(0xFFFFFFFF)).astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32),
                    np.full(nblk, subseq & 0xFFFFFFFF, dtype=np.uint32), np.full(nblk, (subseq >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key)


def randn(n: int, seed: int, subseq: int = 0) -> np.ndarray:
    w = philox_words(n, seed, subseq)
    out = np.empty((w.shape[0], 4), dtype=np.float32)
    for j in range(2):
        u1 = ((w[:, 2 * j] >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
        u2 = ((w[:, 2 * j + 1] >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
        rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
        ang = (np.float32(6.283185307179586) * u2).astype(np.float32)
        out[:, 2 * j] = rad * np.cos(ang)
        out[:, 2 * j + 1] = rad * np.sin(ang)
    return out.reshape(-1)[:n]

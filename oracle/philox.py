"""ORACLE (test infrastructure only) -- counter-based Hutchinson noise shared with the CUDA path.

The reference draws `randn!` from Julia's task-local Xoshiro inside every RHS call
(/root/reference/src/ffjord.jl:92); that stream cannot be reproduced outside Julia, so both
oracle and kernel consume Philox4x32-10 keyed by (seed, stream, rhs_call, trajectory, dim/4)
followed by Box-Muller (documented deviation).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * M0
            p1 = c2.astype(np.uint64) * M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _u01(x):
    # 24-bit mantissa uniform in (0,1): (x>>8 + 0.5) * 2^-24, exact in fp32
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def normal(seed: int, stream: int, call: int, B: int, D: int, b_offset: int = 0) -> np.ndarray:
    """(B, D) float32 standard normals; trajectory index = b_offset + row (global index, so a
    batch shard draws the same numbers as the unsharded run)."""
    nq = (D + 3) // 4
    b = (np.arange(B, dtype=np.uint64) + np.uint64(b_offset)).astype(np.uint32)[:, None]
    q = np.arange(nq, dtype=np.uint32)[None, :]
    x0, x1, x2, x3 = philox4x32_10(b, q, np.uint32(call & 0xFFFFFFFF), np.uint32(stream),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((B, nq, 4), dtype=np.float32)
    for j, (xa, xb) in enumerate(((x0, x1), (x2, x3))):
        u1, u2 = _u01(xa), _u01(xb)
        r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
        th = (np.float32(2.0 * np.pi) * u2).astype(np.float32)
        out[:, :, 2 * j] = r * np.cos(th)
        out[:, :, 2 * j + 1] = r * np.sin(th)
    return out.reshape(B, nq * 4)[:, :D].copy()

"""Philox4x32-10 + Box-Muller, numpy.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Checks the in-register noise generator of the fused Langevin update kernel.  The reference draws its noise
with ``torch.randn_like`` (train.py:326), i.e. from torch's global Philox stream, whose sequence depends on
launch geometry and cannot be reproduced by another kernel; the product therefore keys its own counter-based
stream by (seed, global sample index, step, element quad) so results do not depend on how the batch is
sharded, and parity against the reference is always run with injected noise instead.

Counter layout (must match csrc/langevin_update.cu):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (sample & 0xffffffff, sample >> 32, step, element // 4);  element % 4 picks the lane
Normal transform: u = ((bits >> 8) + 0.5) * 2^-24 in (0,1); (n0, n1) = sqrt(-2 ln u0) * (cos, sin)(2 pi u1);
(n2, n3) likewise from lanes 2, 3.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter [..., 4] uint32, key [..., 2] uint32 (broadcastable) -> [..., 4] uint32."""
    c = [counter[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for r in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        if r < 9:
            k0 = (k0 + np.uint64(W0)) & MASK
            k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def langevin_bits(seed: int, sample0: int, batch: int, nz: int, step: int) -> np.ndarray:
    """uint32 [batch, ceil(nz/4), 4]: the raw Philox output every (sample, quad) of one step consumes."""
    nq = (nz + 3) // 4
    sample = (np.arange(batch, dtype=np.uint64) + np.uint64(sample0))[:, None]
    quad = np.arange(nq, dtype=np.uint64)[None, :]
    ctr = np.stack(np.broadcast_arrays(sample & MASK, sample >> np.uint64(32),
                                       np.full_like(sample, step), quad), axis=-1).astype(np.uint32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key)


def bits_to_normal(bits: np.ndarray) -> np.ndarray:
    """[..., 4] uint32 -> [..., 4] float32 standard normals (Box-Muller on lanes (0,1) and (2,3))."""
    u = ((bits >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    out = np.empty(bits.shape, np.float32)
    for a in (0, 2):
        r = np.sqrt(np.float32(-2.0) * np.log(u[..., a])).astype(np.float32)
        th = (np.float32(2.0) * u[..., a + 1]).astype(np.float64) * np.pi
        out[..., a] = r * np.cos(th).astype(np.float32)
        out[..., a + 1] = r * np.sin(th).astype(np.float32)
    return out


def langevin_noise(seed: int, sample0: int, batch: int, nz: int, step: int) -> np.ndarray:
    """float32 [batch, nz]: the noise the update kernel adds (before the step-size factor)."""
    n = bits_to_normal(langevin_bits(seed, sample0, batch, nz, step))
    return n.reshape(batch, -1)[:, :nz]

"""CPU restatement (numpy) of the library's counter-based Gaussian generator (``pafuse_randn``).  TEST INFRASTRUCTURE ONLY.

The reference draws its sampler noise with ``torch.randn`` / ``randn_like`` (common/diffusionpose.py:283,308); the values
are arbitrary, only their distribution and (for multi-GPU runs) their independence from the sharding matter.  The
library's generator is the published Philox4x32-10 block cipher (Salmon et al., SC'11: multipliers 0xD2511F53 /
0xCD9E8D57, Weyl constants 0x9E3779B9 / 0xBB67AE85) with counter = (element index >> 1, draw number), key = seed, and a
Box-Muller cosine transform in fp64 of words (0,1) for even and (2,3) for odd elements.  This file restates that
definition so that the kernel can be checked value for value on any slice.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter words (uint64 arrays holding 32-bit values); k0/k1 python ints."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2                              # 32x32 -> 64 bit products, exact in uint64
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def randn(seed: int, draw: int, base: int, n: int) -> np.ndarray:
    """fp32 values of global elements base .. base+n-1 of draw ``draw`` under ``seed``."""
    e = np.arange(base, base + n, dtype=np.uint64)
    ctr = e >> np.uint64(1)
    w = philox4x32_10(ctr & MASK, ctr >> np.uint64(32), np.full(n, draw & 0xFFFFFFFF, dtype=np.uint64),
                      np.full(n, (draw >> 32) & 0xFFFFFFFF, dtype=np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    odd = (e & np.uint64(1)).astype(bool)
    a = np.where(odd, w[2], w[0]).astype(np.float64)
    b = np.where(odd, w[3], w[1]).astype(np.float64)
    u1 = (a + 0.5) * (1.0 / 4294967296.0)
    u2 = (b + 0.5) * (1.0 / 4294967296.0)
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)

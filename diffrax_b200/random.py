"""Host-side key management with JAX's threefry semantics (jax.random.key / split).

Used to build the per-trajectory key array ``jr.split(jr.key(seed), N)`` that the reference
vmaps its VirtualBrownianTree over (test/helpers.py:140-169).  NumPy-vectorised restatement of
Threefry-2x32-20 and of jax/_src/prng.py's split layouts.
"""
from __future__ import annotations

import numpy as np

_R = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def threefry2x32(k0, k1, x0, x1):
    """Vectorised block function; all arguments broadcastable uint32 arrays."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, np.uint32)
        k1 = np.asarray(k1, np.uint32)
        ks = (k0, k1, k0 ^ k1 ^ np.uint32(0x1BD11BDA))
        x0 = (np.asarray(x0, np.uint32) + ks[0]).astype(np.uint32)
        x1 = (np.asarray(x1, np.uint32) + ks[1]).astype(np.uint32)
        for g in range(1, 6):
            for r in _R[(g - 1) % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r) ^ x0
            x0 = (x0 + ks[g % 3]).astype(np.uint32)
            x1 = (x1 + ks[(g + 1) % 3] + np.uint32(g)).astype(np.uint32)
    return x0, x1


def key(seed: int) -> np.ndarray:
    """jax.random.key(seed): words (hi32, lo32)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], np.uint32)


def split(k, num: int = 2, *, partitionable: bool = True) -> np.ndarray:
    """jax.random.split(key, num) -> [num, 2] uint32."""
    k = np.asarray(k, np.uint32).reshape(2)
    i = np.arange(num, dtype=np.uint32)
    if partitionable:
        a, b = threefry2x32(k[0], k[1], np.zeros(num, np.uint32), i)
        return np.stack([a, b], axis=1)
    a, b = threefry2x32(k[0], k[1], i, i + np.uint32(num))
    return np.concatenate([a, b]).reshape(num, 2)

"""Shared test helpers: plonky2 binary layouts and seeded field data."""
import struct

import numpy as np

P = 0xFFFFFFFF00000001
GEN = 14293326489335486720

from qpzk.synth import WORMHOLE_GATES, common_bytes  # noqa: E402,F401  (single definition of the layout)


def splitmix64(seed, n):
    """n uniform field elements in [0, p): SplitMix64 with rejection."""
    out = np.empty(n, np.uint64)
    x = seed & 0xFFFFFFFFFFFFFFFF
    i = 0
    M = 0xFFFFFFFFFFFFFFFF
    while i < n:
        x = (x + 0x9E3779B97F4A7C15) & M
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        if z < P:
            out[i] = z
            i += 1
    return out


def rand_felts(rng, shape):
    return rng.integers(0, P, size=shape, dtype=np.uint64)


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r

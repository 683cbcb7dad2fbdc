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


def parse_proof(proof, rows, path_len, fri_steps, final_len, cap_h=4, nq=28, openings=257):
    """`ProofWithPublicInputs::to_bytes` layout, read-only, for the parity tests.

    rows: felts per opened row of the 4 oracles (salted widths), path_len: siblings per oracle path,
    fri_steps: siblings per FRI reduction path (arity 16: 16 ext evals per step), final_len: ext coeffs."""
    a = np.frombuffer(proof, dtype=np.uint8)

    def u64(off, n):
        return np.frombuffer(a[off:off + 8 * n].tobytes(), dtype="<u8").astype(np.uint64)

    ncap = 1 << cap_h
    off = 0
    caps = []
    for _ in range(3):
        caps.append(u64(off, 4 * ncap).reshape(ncap, 4))
        off += 32 * ncap
    opn = u64(off, 2 * openings).reshape(openings, 2)
    off += 16 * openings
    fri_caps = []
    for _ in fri_steps:
        fri_caps.append(u64(off, 4 * ncap).reshape(ncap, 4))
        off += 32 * ncap
    queries = []
    for _ in range(nq):
        q = {"rows": [], "paths": [], "evals": [], "fri_paths": []}
        for w in rows:
            q["rows"].append(u64(off, w))
            off += 8 * w
            assert a[off] == path_len
            q["paths"].append(u64(off + 1, 4 * path_len).reshape(path_len, 4))
            off += 1 + 32 * path_len
        for st in fri_steps:
            q["evals"].append(u64(off, 32))
            off += 256
            assert a[off] == st
            q["fri_paths"].append(u64(off + 1, 4 * st).reshape(st, 4))
            off += 1 + 32 * st
        queries.append(q)
    final = u64(off, 2 * final_len).reshape(final_len, 2)
    off += 16 * final_len
    pow_witness = int(u64(off, 1)[0])
    off += 8
    npi = int(u64(off, 1)[0])
    off += 8
    pis = u64(off, npi)
    off += 8 * npi
    assert off == len(proof)
    return dict(caps=caps, openings=opn, fri_caps=fri_caps, queries=queries, final_poly=final,
                pow_witness=pow_witness, public_inputs=pis)

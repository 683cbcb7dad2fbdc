"""Shared test helpers: plonky2 binary layouts and seeded field data."""
import struct

import numpy as np

P = 0xFFFFFFFF00000001
GEN = 14293326489335486720

GATE_ARITHMETIC, GATE_BASE_SUM_2, GATE_CONSTANT, GATE_NOOP, GATE_POSEIDON, GATE_PUBLIC_INPUT = 0, 2, 3, 9, 11, 12
WORMHOLE_GATES = [(GATE_NOOP, None), (GATE_CONSTANT, 2), (GATE_PUBLIC_INPUT, None), (GATE_BASE_SUM_2, 63),
                  (GATE_ARITHMETIC, 20), (GATE_POSEIDON, None)]


def common_bytes(degree_bits, zk, arities, gates=WORMHOLE_GATES, selector_indices=(0, 0, 0, 0, 0, 1),
                 groups=((0, 5), (5, 6)), num_wires=135, num_routed=80, num_challenges=2, qdf=8,
                 rate_bits=3, cap_height=4, num_queries=28, pow_bits=16, num_gate_constraints=123,
                 num_constants=4, num_public_inputs=16, num_partial_products=9):
    """Serialise a CommonCircuitData the way qp-plonky2 1.1.1 does (layout read off
    /root/reference/wormhole/bench-data/common.bin; reproduced byte-exactly by a test)."""
    u = lambda v: struct.pack("<Q", v)
    fri = u(rate_bits) + u(cap_height) + u(num_queries) + struct.pack("<I", pow_bits) + b"\x01" + u(4) + u(5)
    out = u(num_wires) + u(num_routed) + u(2) + u(100) + u(num_challenges) + u(qdf) + b"\x01" + bytes([1 if zk else 0])
    out += fri + fri + u(len(arities)) + b"".join(u(a) for a in arities) + u(degree_bits) + bytes([1 if zk else 0])
    out += u(len(selector_indices)) + b"".join(u(s) for s in selector_indices)
    out += u(len(groups)) + b"".join(u(a) + u(b) for a, b in groups)
    out += u(qdf) + u(num_gate_constraints) + u(num_constants) + u(num_public_inputs)
    k, ks = 1, []
    for _ in range(num_routed):
        ks.append(k)
        k = k * GEN % P
    out += u(num_routed) + b"".join(u(x) for x in ks)
    out += u(num_partial_products) + u(0) + u(0) + u(0)
    out += u(len(gates))
    for gid, param in gates:
        out += struct.pack("<I", gid) + (u(param) if param is not None else b"")
    return out


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

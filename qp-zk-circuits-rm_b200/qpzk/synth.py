"""Synthetic-witness generator: a tiny "circuit builder" that produces a SATISFYING trace of the
wormhole circuit's shape (SURVEY.md §7 mini-builder recipe): the six gates of
wormhole/bench-data/common.bin (Noop, Constant(2), PublicInput, BaseSum<2>(63), Arithmetic(20),
Poseidon), 135 wires / 80 routed, two selector columns, non-trivial copy-constraint cycles.

Circuit building and witness generation are out of scope for the GPU backend (they stay in Rust),
so tests and the bench need some source of valid (constants, sigmas, wires) triples; this is it
(BASELINE.json: "synthetic witnesses of the bench-data shape"). It is input generation only: the
few hashes it needs come from a `provider` (the GPU library by default, the oracle in CPU tests).
The proofs made from these traces are checked by the oracle's restated plonky2 verifier in tests/.
"""
import ctypes
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


class GpuProvider:
    """Poseidon tables and hashes from libqpzk (the product's own table generator and kernels)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def poseidon_tables(self):
        from . import load_library
        L = load_library()
        L.qpzk_poseidon_tables_host.restype = ctypes.c_void_p
        p = L.qpzk_poseidon_tables_host()
        n = 360 + 12 + 22 + 121 + 242 + 242
        flat = np.frombuffer((ctypes.c_uint64 * n).from_address(p), dtype=np.uint64).copy()
        o = [0, 360, 372, 394, 515, 757, 999]
        return dict(rc=flat[o[0]:o[1]], fast_first=flat[o[1]:o[2]], fast_rc=flat[o[2]:o[3]],
                    fast_init=flat[o[3]:o[4]].reshape(11, 11), fast_w_hat=flat[o[4]:o[5]].reshape(22, 11),
                    fast_v=flat[o[5]:o[6]].reshape(22, 11))

    def hash_no_pad(self, x):
        return self.ctx.hash_no_pad(np.asarray(x, np.uint64))


def root_of_unity(bits):
    r = 7277203076849721926
    for _ in range(32 - bits):
        r = r * r % P
    return r

NOOP, CONSTANT, PUBLIC_INPUT, BASE_SUM, ARITHMETIC, POSEIDON = range(6)
UNUSED = 0xFFFFFFFF
MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]


class PoseidonTracer:
    """Fast-form permutation in Python ints that records the s-box inputs the PoseidonGate wires hold."""

    def __init__(self, provider):
        t = provider.poseidon_tables()
        self.rc = [int(x) for x in t["rc"]]
        self.first = [int(x) for x in t["fast_first"]]
        self.frc = [int(x) for x in t["fast_rc"]]
        self.init = [[int(x) for x in r] for r in t["fast_init"]]
        self.what = [[int(x) for x in r] for r in t["fast_w_hat"]]
        self.v = [[int(x) for x in r] for r in t["fast_v"]]

    @staticmethod
    def mds(s):
        return [(sum(s[(i + r) % 12] * MDS_CIRC[i] for i in range(12)) + (8 * s[0] if r == 0 else 0)) % P
                for r in range(12)]

    def run(self, state):
        s = list(state)
        full0, partial, full1 = [], [], []
        rnd = 0
        for r in range(4):
            s = [(s[i] + self.rc[12 * rnd + i]) % P for i in range(12)]
            if r != 0:
                full0.append(list(s))
            s = self.mds([pow(x, 7, P) for x in s])
            rnd += 1
        s = [(s[i] + self.first[i]) % P for i in range(12)]
        o = [s[0]] + [0] * 11
        for c in range(1, 12):
            o[c] = sum(s[r] * self.init[r - 1][c - 1] for r in range(1, 12)) % P
        s = o
        for r in range(22):
            partial.append(s[0])
            s0 = pow(s[0], 7, P)
            if r < 21:
                s0 = (s0 + self.frc[r]) % P
            d = (s0 * 25 + sum(s[i] * self.what[r][i - 1] for i in range(1, 12))) % P
            s = [d] + [(s[i] + s0 * self.v[r][i - 1]) % P for i in range(1, 12)]
        rnd += 22
        for r in range(4):
            s = [(s[i] + self.rc[12 * rnd + i]) % P for i in range(12)]
            full1.append(list(s))
            s = self.mds([pow(x, 7, P) for x in s])
            rnd += 1
        return s, full0, partial, full1


def build(degree_bits, zk=False, seed=1, mix=(0.46, 0.38, 0.06), arities=None, provider=None):
    """Returns dict(common, digest, constants_sigmas [84][n], wires [135][n], public_inputs [16],
    salts (3 x [4][N]) or None, gate_rows)."""
    if provider is None:
        raise ValueError("a provider (GpuProvider(ctx) or the oracle in CPU tests) is required")
    rng = np.random.default_rng(seed)
    n = 1 << degree_bits
    if arities is None:
        # FriReductionStrategy::ConstantArityBits(4, 5) with rate_bits 3, cap_height 4
        arities, rem = [], degree_bits
        while rem > 5 and rem + 3 - 4 >= 4:
            arities.append(4)
            rem -= 4
    common = common_bytes(degree_bits, zk, arities)
    nw, nr = 135, 80

    wires = rng.integers(0, P, size=(nw, n), dtype=np.uint64).astype(object)
    consts = np.zeros((4, n), dtype=object)
    consts[2] = rng.integers(0, P, size=n, dtype=np.uint64).astype(object)
    consts[3] = rng.integers(0, P, size=n, dtype=np.uint64).astype(object)

    # gate of each row
    gate = np.full(n, NOOP)
    n_bs, n_ar, n_po = int(mix[0] * n), int(mix[1] * n), max(2, int(mix[2] * n))
    rows = rng.permutation(np.arange(1, n))
    gate[0] = PUBLIC_INPUT
    bs_rows = rows[:n_bs]
    ar_rows = rows[n_bs:n_bs + n_ar]
    po_rows = np.sort(rows[n_bs + n_ar:n_bs + n_ar + n_po])
    rest = rows[n_bs + n_ar + n_po:]
    co_rows = rest[: max(1, len(rest) // 2)]
    gate[bs_rows], gate[ar_rows], gate[po_rows], gate[co_rows] = BASE_SUM, ARITHMETIC, POSEIDON, CONSTANT
    consts[0] = np.where(gate == POSEIDON, UNUSED, gate).astype(object)
    consts[1] = np.where(gate == POSEIDON, POSEIDON, UNUSED).astype(object)

    groups = []  # copy-constraint classes: lists of (row, col) that hold equal values

    public_inputs = [int(x) for x in rng.integers(0, P, size=16, dtype=np.uint64)]
    pih = [int(x) for x in provider.hash_no_pad(np.array(public_inputs, np.uint64))]
    for i in range(4):
        wires[i, 0] = pih[i]

    for r in co_rows:
        wires[0, r], wires[1, r] = consts[2, r], consts[3, r]

    for r in bs_rows:
        bits = rng.integers(0, 2, size=63)
        for j in range(63):
            wires[1 + j, r] = int(bits[j])
        wires[0, r] = sum(int(b) << j for j, b in enumerate(bits)) % P

    co_list = list(co_rows)
    for idx, r in enumerate(ar_rows):
        c0, c1 = consts[2, r], consts[3, r]
        share = None
        if co_list and idx % 3 == 0:  # route a Constant gate's output into several addends
            cr = co_list[idx % len(co_list)]
            share = [(cr, 0)]
        for t in range(20):
            if t > 0:  # chain: this op's first multiplicand is the previous op's output
                wires[4 * t, r] = wires[4 * (t - 1) + 3, r]
                groups.append([(r, 4 * (t - 1) + 3), (r, 4 * t)])
            if share is not None and t % 5 == 2:
                wires[4 * t + 2, r] = wires[0, share[0][0]]
                share.append((r, 4 * t + 2))
            m0, m1, ad = wires[4 * t, r], wires[4 * t + 1, r], wires[4 * t + 2, r]
            wires[4 * t + 3, r] = (m0 * m1 % P * c0 + ad * c1) % P
        if share is not None and len(share) > 1:
            groups.append(share)

    tracer = PoseidonTracer(provider)
    prev_out = None
    for r in po_rows:
        if prev_out is not None:  # chain the previous permutation's first 4 outputs into this row's inputs
            for i in range(4):
                wires[i, r] = wires[12 + i, prev_out]
                groups.append([(prev_out, 12 + i), (r, i)])
        swap = int(rng.integers(0, 2))
        wires[24, r] = swap
        inp = [wires[i, r] for i in range(12)]
        st = list(inp)
        for i in range(4):
            delta = swap * (inp[i + 4] - inp[i]) % P
            wires[25 + i, r] = delta
            st[i] = (inp[i] + delta) % P
            st[i + 4] = (inp[i + 4] - delta) % P
        out, full0, partial, full1 = tracer.run(st)
        for rr in range(3):
            for i in range(12):
                wires[29 + 12 * rr + i, r] = full0[rr][i]
        for rr in range(22):
            wires[65 + rr, r] = partial[rr]
        for rr in range(4):
            for i in range(12):
                wires[87 + 12 * rr + i, r] = full1[rr][i]
        for i in range(12):
            wires[12 + i, r] = out[i]
        prev_out = r

    # sigma: identity, then splice the cycles in
    w = root_of_unity(degree_bits)
    omega = [1] * n
    for i in range(1, n):
        omega[i] = omega[i - 1] * w % P
    k_is = [1] * nr
    for j in range(1, nr):
        k_is[j] = k_is[j - 1] * GEN % P
    target = {}
    merged = {}
    for g in groups:  # a cell may appear in several groups: merge them
        cells = []
        for cell in g:
            if cell in merged:
                cells = merged[cell]
                break
        if not cells:
            cells = []
        for cell in g:
            if cell not in cells:
                cells.append(cell)
            merged[cell] = cells
    seen = set()
    for cells in merged.values():
        if id(cells) in seen:
            continue
        seen.add(id(cells))
        vals = {wires[c, r] for r, c in cells}
        assert len(vals) == 1, "copy class with unequal values"
        for a, b in zip(cells, cells[1:] + cells[:1]):
            target[a] = b
    sigmas = np.zeros((nr, n), dtype=object)
    for j in range(nr):
        for i in range(n):
            ti, tj = target.get((i, j), (i, j))
            sigmas[j, i] = k_is[tj] * omega[ti] % P

    cs = np.concatenate([consts, sigmas]).astype(np.uint64)
    N = n << 3
    salts = [rng.integers(0, P, size=(4, N), dtype=np.uint64) for _ in range(3)] if zk else None
    digest = rng.integers(0, P, size=4, dtype=np.uint64)
    return dict(common=common, digest=digest, constants_sigmas=np.ascontiguousarray(cs),
                wires=np.ascontiguousarray(wires.astype(np.uint64)),
                public_inputs=np.array(public_inputs, np.uint64), salts=salts, degree_bits=degree_bits, zk=zk,
                arities=arities, gate=gate)

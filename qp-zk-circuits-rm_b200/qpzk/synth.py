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
        out += struct.pack("<I", gid)
        if param is None:
            continue
        if isinstance(param, (tuple, list)):      # several usize parameters / a length-prefixed field vector
            for x in param:
                out += (u(len(x)) + b"".join(u(int(v)) for v in x)) if isinstance(x, (tuple, list)) else u(x)
        else:
            out += u(param)
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


# ------------------------------------------------------------------------------------------------
# Recursion-shaped circuits (BASELINE configs[4]): the gate set an in-circuit verifier
# (`verify_proof::<C>` at /root/reference/wormhole/aggregator/src/circuits/tree.rs:118-119) instantiates
# under `standard_recursion_config`, with satisfying rows for every gate. The reference ships no
# aggregator circuit data, so the gate MIX here is an estimate and - unlike the wormhole gates - the gate
# definitions themselves are restated from upstream plonky2 without a fixture to pin them.
GATE_ARITHMETIC_EXT, GATE_COSET_INTERPOLATION, GATE_EXPONENTIATION, GATE_MUL_EXT = 1, 4, 5, 8
GATE_POSEIDON_MDS, GATE_RANDOM_ACCESS, GATE_REDUCING_EXT, GATE_REDUCING = 10, 13, 14, 15


def _emul(x, y):
    return ((x[0] * y[0] + 7 * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)


def _eadd(x, y):
    return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)


def _esub(x, y):
    return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)


def _escale(x, s):
    return (x[0] * s % P, x[1] * s % P)


def coset_weights(subgroup_bits):
    """Barycentric weights of the order-2^bits subgroup: w_i = 1 / prod_{j != i} (d_i - d_j)."""
    g = root_of_unity(subgroup_bits)
    dom = [pow(g, i, P) for i in range(1 << subgroup_bits)]
    ws = []
    for i, di in enumerate(dom):
        prod = 1
        for j, dj in enumerate(dom):
            if j != i:
                prod = prod * (di - dj) % P
        ws.append(pow(prod, P - 2, P))
    return dom, ws


# gate list in selector-group order; group sizes + gate degrees stay within the degree-8 quotient bound
R_NOOP, R_CONSTANT, R_PUBLIC_INPUT, R_POSEIDON_MDS, R_BASE_SUM, R_REDUCING, R_REDUCING_EXT, R_ARITHMETIC, \
    R_ARITHMETIC_EXT, R_MUL_EXT, R_EXPONENTIATION, R_RANDOM_ACCESS, R_COSET, R_POSEIDON = range(14)
COSET_BITS, COSET_DEGREE = 4, 6
RECURSION_SELECTORS = (0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 4)
RECURSION_GROUPS = ((0, 6), (6, 10), (10, 12), (12, 13), (13, 14))


def recursion_gates():
    _, ws = coset_weights(COSET_BITS)
    return [(GATE_NOOP, None), (GATE_CONSTANT, 2), (GATE_PUBLIC_INPUT, None), (GATE_POSEIDON_MDS, None),
            (GATE_BASE_SUM_2, 63), (GATE_REDUCING, 43), (GATE_REDUCING_EXT, 32), (GATE_ARITHMETIC, 20),
            (GATE_ARITHMETIC_EXT, 10), (GATE_MUL_EXT, 13), (GATE_EXPONENTIATION, 66),
            (GATE_RANDOM_ACCESS, (4, 4, 2)), (GATE_COSET_INTERPOLATION, (COSET_BITS, COSET_DEGREE, tuple(ws))),
            (GATE_POSEIDON, None)]


def build_recursion(degree_bits, zk=False, seed=1, provider=None, arities=None,
                    mix=(("base_sum", 0.10), ("arithmetic", 0.18), ("arithmetic_ext", 0.16), ("mul_ext", 0.08),
                         ("reducing", 0.06), ("reducing_ext", 0.06), ("random_access", 0.08), ("coset", 0.03),
                         ("exponentiation", 0.03), ("poseidon_mds", 0.02), ("poseidon", 0.14), ("constant", 0.03))):
    """A satisfying trace over the 14-gate recursion set (135 wires / 80 routed, 5 selector columns + 2 gate
    constants). Same return shape as build()."""
    if provider is None:
        raise ValueError("a provider is required")
    rng = np.random.default_rng(seed)
    n = 1 << degree_bits
    if arities is None:
        arities, rem = [], degree_bits
        while rem > 5 and rem + 3 - 4 >= 4:
            arities.append(4)
            rem -= 4
    nsel, nconst = 5, 7
    common = common_bytes(degree_bits, zk, arities, gates=recursion_gates(), selector_indices=RECURSION_SELECTORS,
                          groups=RECURSION_GROUPS, num_constants=nconst)
    nw, nr = 135, 80
    rnd = lambda: int(rng.integers(0, P, dtype=np.uint64))
    rext = lambda: (rnd(), rnd())
    wires = rng.integers(0, P, size=(nw, n), dtype=np.uint64).astype(object)
    consts = np.zeros((nconst, n), dtype=object)
    consts[5] = rng.integers(0, P, size=n, dtype=np.uint64).astype(object)
    consts[6] = rng.integers(0, P, size=n, dtype=np.uint64).astype(object)

    kinds = {"base_sum": R_BASE_SUM, "arithmetic": R_ARITHMETIC, "arithmetic_ext": R_ARITHMETIC_EXT,
             "mul_ext": R_MUL_EXT, "reducing": R_REDUCING, "reducing_ext": R_REDUCING_EXT,
             "random_access": R_RANDOM_ACCESS, "coset": R_COSET, "exponentiation": R_EXPONENTIATION,
             "poseidon_mds": R_POSEIDON_MDS, "poseidon": R_POSEIDON, "constant": R_CONSTANT}
    gate = np.full(n, R_NOOP)
    gate[0] = R_PUBLIC_INPUT
    rows = rng.permutation(np.arange(1, n))
    pos = 0
    for name, frac in mix:
        cnt = max(1, int(frac * n))
        gate[rows[pos:pos + cnt]] = kinds[name]
        pos += cnt
    for sidx, (lo, hi) in enumerate(RECURSION_GROUPS):
        consts[sidx] = np.where((gate >= lo) & (gate < hi), gate, UNUSED).astype(object)

    public_inputs = [int(x) for x in rng.integers(0, P, size=16, dtype=np.uint64)]
    pih = [int(x) for x in provider.hash_no_pad(np.array(public_inputs, np.uint64))]
    for i in range(4):
        wires[i, 0] = pih[i]

    def put_ext(r, w0, v):
        wires[w0, r], wires[w0 + 1, r] = v[0], v[1]

    dom, cw = coset_weights(COSET_BITS)
    nint = ((1 << COSET_BITS) - 2) // (COSET_DEGREE - 1)
    tracer = PoseidonTracer(provider)
    groups = []
    prev_mul_out = None
    for r in range(1, n):
        g = gate[r]
        c0, c1 = consts[5, r], consts[6, r]
        if g == R_CONSTANT:
            wires[0, r], wires[1, r] = c0, c1
        elif g == R_BASE_SUM:
            bits = rng.integers(0, 2, size=63)
            for j in range(63):
                wires[1 + j, r] = int(bits[j])
            wires[0, r] = sum(int(b) << j for j, b in enumerate(bits)) % P
        elif g == R_ARITHMETIC:
            for t in range(20):
                m0, m1, ad = wires[4 * t, r], wires[4 * t + 1, r], wires[4 * t + 2, r]
                wires[4 * t + 3, r] = (m0 * m1 % P * c0 + ad * c1) % P
        elif g == R_ARITHMETIC_EXT:
            for t in range(10):
                w0 = 8 * t
                m0, m1, ad = (wires[w0, r], wires[w0 + 1, r]), (wires[w0 + 2, r], wires[w0 + 3, r]), \
                    (wires[w0 + 4, r], wires[w0 + 5, r])
                put_ext(r, w0 + 6, _eadd(_escale(_emul(m0, m1), c0), _escale(ad, c1)))
        elif g == R_MUL_EXT:
            for t in range(13):
                w0 = 6 * t
                if t == 0 and prev_mul_out is not None:   # route the previous MulExt row's last product in
                    pr = prev_mul_out
                    for d in range(2):
                        wires[d, r] = wires[6 * 12 + 4 + d, pr]
                        groups.append([(pr, 6 * 12 + 4 + d), (r, d)])
                m0, m1 = (wires[w0, r], wires[w0 + 1, r]), (wires[w0 + 2, r], wires[w0 + 3, r])
                put_ext(r, w0 + 4, _escale(_emul(m0, m1), c0))
            prev_mul_out = r
        elif g == R_POSEIDON_MDS:
            ins = [(wires[2 * i, r], wires[2 * i + 1, r]) for i in range(12)]
            for rr in range(12):
                acc = (0, 0)
                for i in range(12):
                    acc = _eadd(acc, _escale(ins[(i + rr) % 12], MDS_CIRC[i]))
                if rr == 0:
                    acc = _eadd(acc, _escale(ins[0], 8))
                put_ext(r, 2 * (12 + rr), acc)
        elif g == R_RANDOM_ACCESS:
            bits, copies, extra, vec = 4, 4, 2, 16
            routed = (2 + vec) * copies + extra
            for cp in range(copies):
                w0 = (2 + vec) * cp
                idx = int(rng.integers(0, vec))
                wires[w0, r] = idx
                wires[w0 + 1, r] = wires[w0 + 2 + idx, r]
                for i in range(bits):
                    wires[routed + cp * bits + i, r] = (idx >> i) & 1
            wires[(2 + vec) * copies, r], wires[(2 + vec) * copies + 1, r] = c0, c1
        elif g in (R_REDUCING, R_REDUCING_EXT):
            ext = g == R_REDUCING_EXT
            ncf = 32 if ext else 43
            cwid = 2 if ext else 1
            start_accs = 6 + ncf * cwid
            alpha, acc = (wires[2, r], wires[3, r]), (wires[4, r], wires[5, r])
            for i in range(ncf):
                cf = (wires[6 + 2 * i, r], wires[7 + 2 * i, r]) if ext else (wires[6 + i, r], 0)
                acc = _eadd(_emul(acc, alpha), cf)
                put_ext(r, 0 if i == ncf - 1 else start_accs + 2 * i, acc)
        elif g == R_EXPONENTIATION:
            nb = 66
            base = wires[0, r]
            bits = [int(b) for b in rng.integers(0, 2, size=nb)]
            for i in range(nb):
                wires[1 + i, r] = bits[i]
            cur = 1
            for i in range(nb):
                prev = 1 if i == 0 else cur * cur % P
                bit = bits[nb - 1 - i]
                cur = prev * (bit * base + (1 - bit)) % P
                wires[2 + nb + i, r] = cur
            wires[1 + nb, r] = cur
        elif g == R_COSET:
            npts = 1 << COSET_BITS
            sp, sv, si = 1 + 2 * npts, 3 + 2 * npts, 5 + 2 * npts
            sshift = si + 4 * nint
            shift = rnd() or 1
            wires[0, r] = shift
            x = (wires[sp, r], wires[sp + 1, r])
            xs = _escale(x, pow(shift, P - 2, P))
            put_ext(r, sshift, xs)
            vals = [(wires[1 + 2 * i, r], wires[2 + 2 * i, r]) for i in range(npts)]

            def partial(lo, hi, ev, pr):
                for i in range(lo, hi):
                    term = _esub(xs, (dom[i], 0))
                    ev = _eadd(_emul(ev, term), _escale(_emul(vals[i], pr), cw[i]))
                    pr = _emul(pr, term)
                return ev, pr

            ev, pr = partial(0, COSET_DEGREE, (0, 0), (1, 0))
            for i in range(nint):
                put_ext(r, si + 2 * i, ev)
                put_ext(r, si + 2 * (nint + i), pr)
                lo = 1 + (COSET_DEGREE - 1) * (i + 1)
                ev, pr = partial(lo, min(lo + COSET_DEGREE - 1, npts), ev, pr)
            put_ext(r, sv, ev)
        elif g == R_POSEIDON:
            swap = int(rng.integers(0, 2))
            wires[24, r] = swap
            inp = [wires[i, r] for i in range(12)]
            st = list(inp)
            for i in range(4):
                delta = swap * (inp[i + 4] - inp[i]) % P
                wires[25 + i, r] = delta
                st[i] = (inp[i] + delta) % P
                st[i + 4] = (inp[i + 4] - delta) % P
            out, full0, partial_in, full1 = tracer.run(st)
            for rr in range(3):
                for i in range(12):
                    wires[29 + 12 * rr + i, r] = full0[rr][i]
            for rr in range(22):
                wires[65 + rr, r] = partial_in[rr]
            for rr in range(4):
                for i in range(12):
                    wires[87 + 12 * rr + i, r] = full1[rr][i]
            for i in range(12):
                wires[12 + i, r] = out[i]

    # sigma: identity with the routed cycles spliced in
    w = root_of_unity(degree_bits)
    omega = np.empty(n, dtype=object)
    acc = 1
    for i in range(n):
        omega[i] = acc
        acc = acc * w % P
    k_is = [1] * nr
    for j in range(1, nr):
        k_is[j] = k_is[j - 1] * GEN % P
    sigmas = np.zeros((nr, n), dtype=object)
    for j in range(nr):
        sigmas[j] = omega * k_is[j] % P
    for cells in groups:
        assert len({wires[c, r] for r, c in cells}) == 1
        for a, b in zip(cells, cells[1:] + cells[:1]):
            sigmas[a[1], a[0]] = k_is[b[1]] * omega[b[0]] % P

    cs = np.concatenate([consts, sigmas]).astype(np.uint64)
    N = n << 3
    salts = [rng.integers(0, P, size=(4, N), dtype=np.uint64) for _ in range(3)] if zk else None
    digest = rng.integers(0, P, size=4, dtype=np.uint64)
    return dict(common=common, digest=digest, constants_sigmas=np.ascontiguousarray(cs),
                wires=np.ascontiguousarray(wires.astype(np.uint64)),
                public_inputs=np.array(public_inputs, np.uint64), salts=salts, degree_bits=degree_bits, zk=zk,
                arities=arities, gate=gate)


def seeded_salts(seed, oracle, n_lde, salt_cols=4):
    """Host restatement of the on-device salt generator (include/qpzk.h QPZK_PROVE_SEEDED_SALTS, csrc/transcript.cuh
    k_salts_chacha8): the [salt_cols][n_lde] blinding columns of oracle `oracle` (0 wires, 1 Z|partial products,
    2 quotient) for a 4-word seed: ChaCha8 blocks with key = seed, counter = block index, nonce = (oracle, 0); each
    block's sixteen words are eight salts; values >= p wrap by p. Vectorised over the blocks."""
    key = np.ascontiguousarray(np.asarray(seed, np.uint64)).view(np.uint32)
    count = salt_cols * n_lde
    nb = (count + 7) // 8
    b = np.arange(nb, dtype=np.uint64)
    s = [np.full(nb, v, np.uint32) for v in (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)]
    s += [np.full(nb, key[i], np.uint32) for i in range(8)]
    s += [(b & np.uint64(0xFFFFFFFF)).astype(np.uint32), (b >> np.uint64(32)).astype(np.uint32),
          np.full(nb, oracle, np.uint32), np.zeros(nb, np.uint32)]
    x = [v.copy() for v in s]

    def rotl(v, n):
        return (v << np.uint32(n)) | (v >> np.uint32(32 - n))

    def qr(a, bb, c, d):
        x[a] += x[bb]; x[d] = rotl(x[d] ^ x[a], 16)
        x[c] += x[d]; x[bb] = rotl(x[bb] ^ x[c], 12)
        x[a] += x[bb]; x[d] = rotl(x[d] ^ x[a], 8)
        x[c] += x[d]; x[bb] = rotl(x[bb] ^ x[c], 7)

    with np.errstate(over="ignore"):
        for _ in range(4):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        w = [x[i] + s[i] for i in range(16)]
    out = np.empty((nb, 8), np.uint64)
    for e in range(8):
        v = (w[2 * e + 1].astype(np.uint64) << np.uint64(32)) | w[2 * e].astype(np.uint64)
        out[:, e] = np.where(v >= np.uint64(P), v - np.uint64(P), v)
    return out.ravel()[:count].reshape(salt_cols, n_lde)

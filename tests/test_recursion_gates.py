"""The recursion gate set (BASELINE configs[4]): ArithmeticExtension, MulExtension,
PoseidonMds, RandomAccess, Reducing, ReducingExtension, Exponentiation, CosetInterpolation next to the
six wormhole gates - what an in-circuit `verify_proof::<C>` instantiates
(/root/reference/wormhole/aggregator/src/circuits/tree.rs:111-127).

Honest status: qp-plonky2 is un-vendored and the reference ships no aggregator circuit data, so these
gate definitions are restated from upstream plonky2 WITHOUT a fixture to pin them ("parity unpinned").
What is checked: (CPU) proofs over a satisfying recursion-shaped trace are accepted by the restated
verifier, which evaluates the same constraints over F_p^2 at zeta, and one flipped bit in a row of EACH
gate type is rejected; (GPU) `qpzk_prove` emits the very bytes the oracle prover does."""
import numpy as np
import pytest

import minibuilder
from oracle import oracle as orc
from qpzk import synth

# (gate index in the recursion list, a wire column one of its constraints reads)
CONSTRAINED_CELL = {
    synth.R_ARITHMETIC_EXT: 6, synth.R_MUL_EXT: 5, synth.R_POSEIDON_MDS: 30, synth.R_RANDOM_ACCESS: 1,
    synth.R_REDUCING: 0, synth.R_REDUCING_EXT: 80, synth.R_EXPONENTIATION: 70, synth.R_COSET: 36,
    synth.R_POSEIDON: 70, synth.R_BASE_SUM: 0, synth.R_ARITHMETIC: 3, synth.R_CONSTANT: 1,
}


def test_recursion_common_roundtrip_and_sizes():
    circ = minibuilder.build_recursion(7, seed=3)
    assert circ["constants_sigmas"].shape == (5 + 2 + 80, 128)          # 5 selector columns + 2 gate constants
    assert set(np.unique(circ["gate"])) == set(range(14))               # every gate type has rows


@pytest.mark.parametrize("k,zk", [(7, False), (8, True)])
def test_oracle_proves_and_verifies_recursion_shape(k, zk):
    circ = minibuilder.build_recursion(k, zk=zk, seed=10 + k)
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=8)
    proof = oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    rc, _ = orc.verify(circ["common"], oc.verifier_only_bytes(), proof)
    assert rc == 0
    assert orc.proof_roundtrip(circ["common"], proof) == 1


def test_every_gate_type_is_constrained():
    circ = minibuilder.build_recursion(8, zk=False, seed=5)
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=8)
    for g, col in CONSTRAINED_CELL.items():
        rows = np.where(circ["gate"] == g)[0]
        assert len(rows) > 0
        w = circ["wires"].copy()
        w[col, rows[0]] ^= np.uint64(1)
        proof = oc.prove(w, circ["public_inputs"], circ["salts"])
        rc, _ = orc.verify(circ["common"], oc.verifier_only_bytes(), proof)
        assert rc != 0, "flipping wire %d of a row of gate %d went unnoticed" % (col, g)


@pytest.mark.gpu
@pytest.mark.parametrize("k,zk", [(7, False), (9, True), (12, False), (15, True)])
def test_gpu_recursion_proof_matches_oracle(k, zk):
    import qpzk
    ctx = qpzk.Context(0)
    circ = minibuilder.build_recursion(k, zk=zk, seed=30 + k)
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=16)
    want = oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    got = gc.prove(circ["wires"], circ["public_inputs"], circ["salts"], trace=True)
    tr = oc.trace(rounds=len(circ["arities"]))
    n = 1 << k
    assert np.array_equal(gc.trace(1).reshape(-1, n), tr["zs_pp"])
    assert np.array_equal(gc.trace(2).reshape(-1, n), tr["quotient_chunks"])     # every gate's constraints, on the LDE coset
    assert got == want
    rc, _ = orc.verify(circ["common"], gc.verifier_only_bytes(), got)
    assert rc == 0
    # a witness that violates one recursion gate: the GPU proof must be rejected too
    w = circ["wires"].copy()
    w[36, int(np.where(circ["gate"] == synth.R_COSET)[0][0])] ^= np.uint64(1)
    rc, _ = orc.verify(circ["common"], gc.verifier_only_bytes(), gc.prove(w, circ["public_inputs"], circ["salts"]))
    assert rc != 0
    gc.free()
    ctx.close()


@pytest.mark.gpu
def test_gpu_recursion_proof_of_2p18_rows_is_accepted():
    """BASELINE configs[4] quotes aggregation nodes of degree ~2^17-2^18: the largest of them through
    `qpzk_prove` (a 2^21-point quotient domain, the two-pass IFFT beyond 2^20 points) and the restated verifier.
    Too large for the CPU oracle prover in a test, so acceptance and tamper-rejection are the checks."""
    import qpzk
    ctx = qpzk.Context(0)
    circ = synth.build_recursion(18, zk=True, seed=48, provider=synth.GpuProvider(ctx))
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    proof = gc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    rc, _ = orc.verify(circ["common"], gc.verifier_only_bytes(), proof)
    assert rc == 0
    w = circ["wires"].copy()
    w[36, int(np.where(circ["gate"] == synth.R_COSET)[0][-1])] ^= np.uint64(1)
    rc, _ = orc.verify(circ["common"], gc.verifier_only_bytes(), gc.prove(w, circ["public_inputs"], circ["salts"]))
    assert rc != 0
    gc.free()
    ctx.close()


# ---- what the gates MEAN (independent of wire order): the rows the builder fills, and the constraints accept,
# compute the functions an in-circuit FRI verifier needs them for ----
def _ext(w, r, i):
    return (int(w[i, r]), int(w[i + 1, r]))


def _emul(x, y):
    P = orc.P
    return ((x[0] * y[0] + 7 * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)


def _einv(x):
    P = orc.P
    d = (x[0] * x[0] - 7 * x[1] * x[1]) % P
    di = pow(d, P - 2, P)
    return (x[0] * di % P, (-x[1]) * di % P)


def test_gate_semantics():
    P = orc.P
    circ = minibuilder.build_recursion(8, zk=False, seed=77)
    w, gate = circ["wires"], circ["gate"]
    # CosetInterpolation: evaluation_value == the degree-<16 interpolant of (shift*g^i, value_i) at evaluation_point
    r = int(np.where(gate == synth.R_COSET)[0][0])
    shift = int(w[0, r])
    g = orc.root_of_unity(4)
    xs = [shift * pow(g, i, P) % P for i in range(16)]
    vals = [_ext(w, r, 1 + 2 * i) for i in range(16)]
    x = _ext(w, r, 33)
    acc = (0, 0)
    for i in range(16):   # Lagrange, in F_p^2
        num, den = (1, 0), 1
        for j in range(16):
            if j != i:
                num = _emul(num, ((x[0] - xs[j]) % P, x[1]))
                den = den * (xs[i] - xs[j]) % P
        term = _emul(vals[i], num)
        di = pow(den, P - 2, P)
        acc = ((acc[0] + term[0] * di) % P, (acc[1] + term[1] * di) % P)
    assert acc == _ext(w, r, 35)
    # Reducing: output == old_acc*alpha^n + sum coeff_i * alpha^(n-1-i)   (Horner)
    r = int(np.where(gate == synth.R_REDUCING)[0][0])
    alpha, acc = _ext(w, r, 2), _ext(w, r, 4)
    for i in range(43):
        acc = _emul(acc, alpha)
        acc = ((acc[0] + int(w[6 + i, r])) % P, acc[1])
    assert acc == _ext(w, r, 0)
    # Exponentiation: output == base ^ (little-endian power bits)
    r = int(np.where(gate == synth.R_EXPONENTIATION)[0][0])
    e = sum(int(w[1 + i, r]) << i for i in range(66))
    assert pow(int(w[0, r]), e, P) == int(w[67, r])
    # RandomAccess: claimed_element == list[access_index] for every copy
    r = int(np.where(gate == synth.R_RANDOM_ACCESS)[0][0])
    for cp in range(4):
        assert int(w[18 * cp + 1, r]) == int(w[18 * cp + 2 + int(w[18 * cp, r]), r])
    # MulExtension / ArithmeticExtension: field arithmetic in F_p^2
    r = int(np.where(gate == synth.R_MUL_EXT)[0][0])
    c0 = int(circ["constants_sigmas"][5, r])
    m = _emul(_ext(w, r, 0), _ext(w, r, 2))
    assert (m[0] * c0 % P, m[1] * c0 % P) == _ext(w, r, 4)
    # and F_p^2 really is a field with X^2 = 7: x * x^-1 == 1
    assert _emul(x, _einv(x)) == (1, 0)

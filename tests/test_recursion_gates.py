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
@pytest.mark.parametrize("k,zk", [(7, False), (9, True), (12, False)])
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

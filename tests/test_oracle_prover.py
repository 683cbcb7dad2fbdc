"""The restated `prove()` (oracle/prover.hpp) on satisfying synthetic wormhole-shaped traces: its
proofs must be accepted by the restated verifier (which is pinned by the reference's shipped proof),
tampered witnesses / proofs must be rejected, and the byte layout must match the shipped sizes."""
import numpy as np
import pytest

import minibuilder
from oracle import oracle as orc


@pytest.fixture(scope="module")
def small():
    return minibuilder.build(6, zk=False, seed=3)


def _prove(circ, threads=8):
    c = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=threads)
    proof = c.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    return c, proof


def test_small_proof_verifies(small):
    c, proof = _prove(small)
    rc, ch = orc.verify(small["common"], c.verifier_only_bytes(), proof)
    assert rc == 0
    assert orc.proof_roundtrip(small["common"], proof) == 1
    # deterministic: same inputs, same bytes
    assert c.prove(small["wires"], small["public_inputs"]) == proof


def test_zk_shape_proof_verifies():
    circ = minibuilder.build(7, zk=True, seed=4)
    c, proof = _prove(circ)
    rc, _ = orc.verify(circ["common"], c.verifier_only_bytes(), proof)
    assert rc == 0


def test_bad_witness_is_not_provable_or_rejected(small):
    """A wire that breaks a gate constraint makes the vanishing polynomial indivisible by Z_H: the
    resulting "proof" must be rejected (upstream panics with 'not divisible'; here the quotient's
    high coefficients are simply non-zero and the verifier's identity fails)."""
    bad = dict(small)
    w = small["wires"].copy()
    row = int(np.where(small["gate"] == minibuilder.ARITHMETIC)[0][0])
    w[3, row] ^= np.uint64(1)   # corrupt an arithmetic output
    bad["wires"] = w
    c, proof = _prove(bad)
    rc, _ = orc.verify(bad["common"], c.verifier_only_bytes(), proof)
    assert rc != 0


def test_broken_copy_constraint_is_rejected(small):
    bad = dict(small)
    w = small["wires"].copy()
    row = int(np.where(small["gate"] == minibuilder.ARITHMETIC)[0][0])
    # break the chain out(op0) -> m0(op1) while keeping both gates locally satisfied
    w[4, row] = (int(w[4, row]) + 1) % orc.P
    c0, c1 = int(small["constants_sigmas"][2, row]), int(small["constants_sigmas"][3, row])
    w[7, row] = (int(w[4, row]) * int(w[5, row]) % orc.P * c0 + int(w[6, row]) * c1) % orc.P
    # (later ops in the chain are now inconsistent too, which is fine: still must be rejected)
    bad["wires"] = w
    c, proof = _prove(bad)
    rc, _ = orc.verify(bad["common"], c.verifier_only_bytes(), proof)
    assert rc != 0


def test_tampered_public_input_is_rejected(small):
    c, proof = _prove(small)
    vo = c.verifier_only_bytes()
    for lane in range(0, 16 * 8, 9):  # verifier_tests.rs:48-66 flips every PI byte lane
        bad = bytearray(proof)
        bad[len(proof) - 128 + lane] ^= 1
        try:
            rc, _ = orc.verify(small["common"], vo, bytes(bad))
        except RuntimeError:
            rc = -1
        assert rc != 0


@pytest.mark.parametrize("k,zk,size", [(13, False, 132712), (14, True, 148932)])
def test_wormhole_shapes_sizes_and_acceptance(k, zk, size):
    """Same shapes as the shipped proofs: serialised sizes must match exactly."""
    circ = minibuilder.build(k, zk=zk, seed=5)
    c, proof = _prove(circ, threads=8)
    assert len(proof) == size
    rc, _ = orc.verify(circ["common"], c.verifier_only_bytes(), proof)
    assert rc == 0

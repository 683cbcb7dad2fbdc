"""GPU prover parity: `qpzk_prove` (the device path behind `ProverCircuitData::prove`) against the
oracle's restated prover on the same synthetic wormhole-shaped circuits and the same injected salts:
every stage's intermediate values, then the final ProofWithPublicInputs bytes, must be identical,
and the bytes must be accepted by the restated verifier."""
import numpy as np
import pytest

import minibuilder
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import qpzk
    c = qpzk.Context(0)
    yield c
    c.close()


def _both(ctx, circ, threads=16):
    import qpzk
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=threads)
    want = oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    got = gc.prove(circ["wires"], circ["public_inputs"], circ["salts"], trace=True)
    return oc, want, gc, got


@pytest.mark.parametrize("k,zk", [(6, False), (7, True), (9, False), (13, False), (14, True)])
def test_prove_matches_oracle_stage_by_stage(ctx, k, zk):
    circ = minibuilder.build(k, zk=zk, seed=20 + k)
    oc, want, gc, got = _both(ctx, circ)
    assert gc.verifier_only_bytes() == oc.verifier_only_bytes()           # constants|sigmas cap (build())
    tr = oc.trace(rounds=len(circ["arities"]))
    n = 1 << k
    ch = gc.trace(0)
    want_ch = np.concatenate([tr["betas"], tr["gammas"], tr["alphas"], tr["zeta"], tr["fri_alpha"],
                              tr["fri_betas"].ravel()])
    assert np.array_equal(gc.trace(1).reshape(-1, n), tr["zs_pp"])       # H8 Z / partial products
    assert np.array_equal(ch[:6], want_ch[:6])                            # betas, gammas, alphas
    assert np.array_equal(gc.trace(2).reshape(-1, n), tr["quotient_chunks"])   # H9 quotient chunks
    assert np.array_equal(ch[6:8], want_ch[6:8])                          # zeta
    assert np.array_equal(ch[8:10], want_ch[8:10])                        # FRI alpha (after openings, H10)
    assert np.array_equal(gc.trace(3).reshape(n, 2), tr["final_poly"])   # H11 polynomial entering FRI
    assert np.array_equal(ch, want_ch)                                    # H12 betas
    assert got == want                                                    # H13/H14/H16: identical proof bytes
    rc, _ = orc.verify(circ["common"], gc.verifier_only_bytes(), got)
    assert rc == 0
    gc.free()


def test_wormhole_shape_sizes(ctx):
    import qpzk
    for k, zk, size in ((13, False, 132712), (14, True, 148932)):
        circ = minibuilder.build(k, zk=zk, seed=7)
        gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
        proof = gc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
        assert len(proof) == size
        rc, _ = orc.verify(circ["common"], gc.verifier_only_bytes(), proof)
        assert rc == 0
        # second proof on the same circuit object (device-resident constants|sigmas batch is reused)
        assert gc.prove(circ["wires"], circ["public_inputs"], circ["salts"]) == proof
        gc.free()


def test_invalid_witness_is_rejected_by_verifier(ctx):
    import qpzk
    circ = minibuilder.build(8, zk=False, seed=9)
    w = circ["wires"].copy()
    row = int(np.where(circ["gate"] == minibuilder.POSEIDON)[0][0])
    w[70, row] ^= np.uint64(1)      # corrupt one partial-round s-box wire
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    proof = gc.prove(w, circ["public_inputs"])
    rc, _ = orc.verify(circ["common"], gc.verifier_only_bytes(), proof)
    assert rc != 0
    gc.free()


def test_unsupported_gate_set_is_refused(ctx):
    import qpzk
    from helpers import common_bytes
    bad = common_bytes(6, False, [4], gates=[(9, None), (3, 2), (12, None), (2, 63), (0, 20), (27, None)])
    with pytest.raises(qpzk.QpzkError):
        qpzk.Circuit(ctx, bad, np.zeros(4, np.uint64), np.zeros((84, 64), np.uint64))


@pytest.mark.parametrize("k,zk", [(8, False), (12, True)])
def test_stage_hooks_match_the_oracle_trace(ctx, k, zk):
    """qpzk_zs_partial_products / qpzk_quotient / qpzk_batch_eval_ext called one by one - the way a qp-plonky2
    fork that keeps its own prove() loop would - with the challenges of the oracle's transcript: each
    stage must reproduce the oracle prover's intermediate values exactly."""
    import qpzk
    circ = minibuilder.build(k, zk=zk, seed=50 + k)
    n = 1 << k
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=8)
    oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    tr = oc.trace(rounds=len(circ["arities"]))
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    salts = circ["salts"] or [None, None, None]
    # H8
    zs = gc.zs_partial_products(circ["wires"], tr["betas"], tr["gammas"])
    assert np.array_equal(zs, tr["zs_pp"])
    # commits (H2-H7), then H9 on the committed batches
    wb = qpzk.PolynomialBatch.from_values(ctx, circ["wires"], 3, 4, salts=salts[0])
    zb = qpzk.PolynomialBatch.from_values(ctx, zs, 3, 4, salts=salts[1])
    pih = orc.hash_no_pad(circ["public_inputs"])
    chunks = gc.quotient(wb, zb, pih, tr["betas"], tr["gammas"], tr["alphas"])
    assert np.array_equal(chunks, tr["quotient_chunks"])
    # H10: openings of the quotient oracle at zeta
    qb = qpzk.PolynomialBatch.from_coeffs(ctx, chunks, 3, 4, salts=salts[2])
    ev = qb.eval_ext(tr["zeta"])
    z = (int(tr["zeta"][0]), int(tr["zeta"][1]))
    P = orc.P
    for c in (0, 15):
        a0, a1 = 0, 0
        for coef in chunks[c][::-1]:
            a0, a1 = (a0 * z[0] + 7 * a1 * z[1] + int(coef)) % P, (a0 * z[1] + a1 * z[0]) % P
        assert (int(ev[c, 0]), int(ev[c, 1])) == (a0, a1)
    # shape checks are enforced
    with pytest.raises(qpzk.QpzkError):
        gc.quotient(zb, wb, pih, tr["betas"], tr["gammas"], tr["alphas"])
    for x in (wb, zb, qb):
        x.free()
    gc.free()


@pytest.mark.parametrize("k,zk", [(9, False), (13, True)])
def test_fri_object_reproduces_the_oracle_proof(ctx, k, zk):
    """The FRI prover driven step by step through the C ABI (begin / commit_round / fold / final_poly / pow /
    query) with the oracle transcript's challenges: commit-phase caps, final polynomial, proof-of-work
    witness and every query opening must be the ones inside the oracle prover's proof bytes."""
    import qpzk
    from helpers import parse_proof
    circ = minibuilder.build(k, zk=zk, seed=70 + k)
    n, salt = 1 << k, 4 if zk else 0
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=8)
    proof = oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    tr = oc.trace(rounds=len(circ["arities"]))
    rc, chal = orc.verify(circ["common"], oc.verifier_only_bytes(), proof)
    assert rc == 0
    steps = []
    lg = k + 3
    for a in circ["arities"]:
        lg -= a
        steps.append(lg - 4)
    final_len = 1 << (k - sum(circ["arities"]))
    pr = parse_proof(proof, rows=[84, 135 + salt, 20 + salt, 16 + salt], path_len=k + 3 - 4, fri_steps=steps,
                     final_len=final_len)
    salts = circ["salts"] or [None, None, None]
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    wb = qpzk.PolynomialBatch.from_values(ctx, circ["wires"], 3, 4, salts=salts[0])
    zb = qpzk.PolynomialBatch.from_values(ctx, tr["zs_pp"], 3, 4, salts=salts[1])
    qb = qpzk.PolynomialBatch.from_coeffs(ctx, tr["quotient_chunks"], 3, 4, salts=salts[2])
    assert np.array_equal(wb.cap, pr["caps"][0]) and np.array_equal(zb.cap, pr["caps"][1])
    assert np.array_equal(qb.cap, pr["caps"][2])
    f = qpzk.Fri(gc, wb, zb, qb, tr["zeta"], tr["fri_alpha"])
    assert f.num_rounds == len(circ["arities"])
    betas = tr["fri_betas"].reshape(-1, 2)
    for rnd in range(f.num_rounds):
        assert np.array_equal(f.commit_round(), pr["fri_caps"][rnd])
        f.fold(betas[rnd])
    assert np.array_equal(f.final_poly(), pr["final_poly"])
    # every query round of the proof, opened one index at a time
    widths = [84, 135 + salt, 20 + salt, 16 + salt]
    L0 = k + 3 - 4
    for qi, x in enumerate(int(v) for v in chal["query_indices"][:28]):
        out = f.query(x)
        off = 0
        for o in range(4):
            assert np.array_equal(out[off:off + widths[o]], pr["queries"][qi]["rows"][o])
            off += widths[o]
            assert np.array_equal(out[off:off + 4 * L0].reshape(L0, 4), pr["queries"][qi]["paths"][o])
            off += 4 * L0
        for s_, st in enumerate(steps):
            assert np.array_equal(out[off:off + 32], pr["queries"][qi]["evals"][s_])
            off += 32
            assert np.array_equal(out[off:off + 4 * st].reshape(st, 4), pr["queries"][qi]["fri_paths"][s_])
            off += 4 * st
        assert off == out.size
    # calling the steps out of order is refused
    with pytest.raises(qpzk.QpzkError):
        f.fold(betas[0])
    f.free()
    for x in (wb, zb, qb):
        x.free()
    gc.free()

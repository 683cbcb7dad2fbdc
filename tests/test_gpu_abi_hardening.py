"""The C ABI under hostile or sloppy callers: untrusted CommonCircuitData is range-checked before
anything is launched, element counts are verified, shards refuse leaves they do not hold, and one host thread
can keep several contexts busy through qpzk_prove_begin / qpzk_prove_end."""
import numpy as np
import pytest

import minibuilder
from helpers import rand_felts
from oracle import oracle as orc
from qpzk import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import qpzk
    c = qpzk.Context(0)
    yield c
    c.close()


def _bad_common(**kw):
    k = kw.pop("k", 6)
    args = dict(degree_bits=k, zk=False, arities=[4])
    args.update(kw)
    return synth.common_bytes(**args)


@pytest.mark.parametrize("kw", [
    dict(selector_indices=(0, 0, 0, 0, 0, 7)),              # selector index outside the groups
    dict(groups=((0, 5), (5, 9))),                          # group end beyond the gate list
    dict(k=40),                                             # degree_bits: 1 << 40 rows
    dict(rate_bits=40),
    dict(cap_height=60),
    dict(num_wires=100),                                    # PoseidonGate needs 135 wires
    dict(num_constants=3),                                  # ArithmeticGate needs 2 constants after 2 selectors
    dict(num_partial_products=3),                           # 4 chunks of 8 < 80 routed wires
    dict(num_gate_constraints=50),                          # PoseidonGate emits 123
    dict(num_queries=100000),
    dict(arities=[4, 4, 4]),                                # folds below the cap at 2^6 rows
    dict(num_challenges=9),
    dict(pow_bits=99),
    dict(qdf=6),
])
def test_malformed_common_data_is_refused(ctx, kw):
    import qpzk
    cb = _bad_common(**kw)
    cs = np.zeros((84, 64), np.uint64)
    with pytest.raises(qpzk.QpzkError) as e:
        qpzk.Circuit(ctx, cb, np.zeros(4, np.uint64), cs)
    assert e.value.code in (-1, -5)
    # truncated bytes, too
    with pytest.raises(qpzk.QpzkError):
        qpzk.Circuit(ctx, _bad_common()[:100], np.zeros(4, np.uint64), cs)


def test_element_counts_are_checked(ctx):
    import qpzk
    circ = minibuilder.build(6, zk=True, seed=3)
    with pytest.raises(qpzk.QpzkError):                      # constants|sigmas one column short
        qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"][:-1])
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    assert gc.constants_sigmas_cap.shape == (16, 4)
    with pytest.raises(qpzk.QpzkError):                      # witness one column short
        gc.prove(circ["wires"][:-1], circ["public_inputs"], circ["salts"])
    with pytest.raises(qpzk.QpzkError):                      # salts of the wrong length
        gc.prove(circ["wires"], circ["public_inputs"], [s[:, :-8] for s in circ["salts"]])
    with pytest.raises(qpzk.QpzkError):                      # hiding circuit without salts
        gc.prove(circ["wires"], circ["public_inputs"], None)
    with pytest.raises(qpzk.QpzkError):                      # wrong number of public inputs
        gc.prove(circ["wires"], circ["public_inputs"][:-1], circ["salts"])
    with pytest.raises(qpzk.QpzkError):                      # nothing in flight
        gc.prove_end()
    # and the handle still works afterwards
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=4)
    assert gc.prove(circ["wires"], circ["public_inputs"], circ["salts"]) == oc.prove(circ["wires"], circ["public_inputs"],
                                                                                     circ["salts"])
    gc.free()


def test_shard_refuses_foreign_leaves(ctx):
    import qpzk
    rng = np.random.default_rng(5)
    k, ncols, r, cap_h = 8, 9, 3, 4
    n, N = 1 << k, 1 << (k + r)
    vals = rand_felts(rng, (ncols, n))
    want = orc.batch_commit(vals, r, cap_h, threads=4)
    d = ctx.dev_alloc(vals.nbytes)
    ctx.h2d(d, vals)
    sh = qpzk.PolynomialBatch.from_values_shard_dev(ctx, d, ncols, n, r, cap_h, 4, 8)   # leaves [N/4, N/2)
    row, sib = sh.open(N // 4)
    assert np.array_equal(row, want["leaves"][N // 4])
    for leaf in (0, N // 4 - 1, N // 2, N - 1):
        with pytest.raises(qpzk.QpzkError):
            sh.open(leaf)
    # get_lde_values: natural index i lives at leaf bitrev(i); index 0 -> leaf 0 is foreign, index 2 -> leaf N/4 is owned
    assert np.array_equal(sh.get_lde_values(2), want["leaves"][N // 4][:ncols])
    with pytest.raises(qpzk.QpzkError):
        sh.get_lde_values(0)
    with pytest.raises(qpzk.QpzkError):
        sh.export()
    sh.free()
    ctx.dev_free(d)


def test_one_thread_drives_several_contexts():
    """qpzk_prove_begin on four contexts back to back, then qpzk_prove_end on each: four proofs in flight from
    a single host thread, every one byte-identical to the oracle's."""
    import qpzk
    ctxs = [qpzk.Context(0) for _ in range(4)]
    circs = [minibuilder.build(7 + (i & 1), zk=bool(i & 1), seed=40 + i) for i in range(4)]
    gcs = [qpzk.Circuit(c, z["common"], z["digest"], z["constants_sigmas"]) for c, z in zip(ctxs, circs)]
    for rep in range(3):
        for gc, z in zip(gcs, circs):
            gc.prove_begin(z["wires"], z["public_inputs"], z["salts"])
        with pytest.raises(qpzk.QpzkError):                  # one proof in flight per handle
            gcs[0].prove_begin(circs[0]["wires"], circs[0]["public_inputs"], circs[0]["salts"])
        proofs = [gc.prove_end() for gc in gcs]
        for z, p in zip(circs, proofs):
            oc = orc.Circuit(z["common"], z["digest"], z["constants_sigmas"], threads=4)
            assert p == oc.prove(z["wires"], z["public_inputs"], z["salts"])
    for gc in gcs:
        gc.free()
    for c in ctxs:
        c.close()


def test_small_and_ragged_trees_through_the_climb_kernel(ctx):
    """MerkleTree::new shapes that exercise every branch of the one-launch tree climb: leaves in the cap,
    a single level, short rows (hash_or_noop copies), counts around the 16-lane threshold."""
    import qpzk
    rng = np.random.default_rng(9)
    for log_n, leaf_len, cap_h in ((0, 3, 0), (1, 9, 0), (1, 9, 1), (4, 4, 4), (4, 4, 0), (5, 135, 2), (12, 20, 4),
                                   (13, 16, 4), (14, 5, 0), (3, 1, 1)):
        leaves = rand_felts(rng, (1 << log_n, leaf_len))
        t = qpzk.MerkleTree(ctx, leaves, cap_h)
        dg, cap = orc.merkle_new(leaves, cap_h, threads=4)
        assert np.array_equal(t.cap, cap), (log_n, leaf_len, cap_h)
        assert np.array_equal(t.digests, dg), (log_n, leaf_len, cap_h)
        t.free()


def test_seeded_salts_match_the_host_restatement(ctx):
    """QPZK_PROVE_SEEDED_SALTS: the blinding salts are drawn on the device from a 32-byte seed (ChaCha8). The proof
    must be the one the oracle prover produces when it is handed the same salts, restated on the host."""
    import qpzk
    circ = minibuilder.build(8, zk=True, seed=31)
    seed = np.array([0x0123456789ABCDEF, 0xFEDCBA9876543210, 7, 0xFFFFFFFFFFFFFFFF], np.uint64)
    N = 1 << (8 + 3)
    salts = [synth.seeded_salts(seed, o, N) for o in range(3)]
    assert all((s < np.uint64(0xFFFFFFFF00000001)).all() for s in salts)
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=4)
    want = oc.prove(circ["wires"], circ["public_inputs"], salts)
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    assert gc.prove(circ["wires"], circ["public_inputs"], seed=seed) == want
    assert gc.prove(circ["wires"], circ["public_inputs"], salts) == want          # explicit salts: same bytes
    other = gc.prove(circ["wires"], circ["public_inputs"], seed=seed + np.uint64(1))
    assert other != want                                                           # another seed, another proof
    rc, _ = orc.verify(circ["common"], gc.verifier_only_bytes(), other)
    assert rc == 0
    with pytest.raises(ValueError):                                                # a seed is 4 words
        gc.prove(circ["wires"], circ["public_inputs"], seed=seed[:3])
    gc.free()

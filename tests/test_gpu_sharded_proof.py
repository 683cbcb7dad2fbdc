"""ONE proof over several GPUs (BASELINE configs[4]): every rank runs the phases of
qpzk_sprove_* on its own range of cap subtrees and the ranks exchange subtree roots, quotient values and opened
rows in between. Here the ranks are separate contexts on cuda:0 and the exchanges go through the host
(qpzk.dist.exchange_local), which exercises the whole sharding logic on one GPU; the NCCL spelling
(exchange_nccl) moves the same buffers and is run by bench.py at N > 1. Every rank must end with the bytes of
the single-GPU proof, which are the oracle's."""
import numpy as np
import pytest

import minibuilder
from oracle import oracle as orc
from qpzk import dist as qdist
from qpzk import synth

pytestmark = pytest.mark.gpu


def _prove_sharded(circ, world, cap_h=4, rate_bits=3, on_device=False):
    import qpzk
    ctxs = [qpzk.Context(0) for _ in range(world)]
    gcs = [qpzk.Circuit(c, circ["common"], circ["digest"], circ["constants_sigmas"]) for c in ctxs]
    wires = [circ["wires"]] * world
    salts = [circ["salts"]] * world
    dev = []
    if on_device:
        wires, salts = [], []
        for c in ctxs:
            d = c.dev_alloc(circ["wires"].nbytes)
            c.h2d(d, circ["wires"])
            ds = None
            if circ["salts"] is not None:
                ds = []
                for s in circ["salts"]:
                    p = c.dev_alloc(s.nbytes)
                    c.h2d(p, s)
                    ds.append(p)
            wires.append(d)
            salts.append(ds)
            dev.append((c, [d] + (ds or [])))
    proofs = []
    for rank, gc in enumerate(gcs):
        b, e = qdist.shard_subtrees(rank, world, cap_h, rate_bits)
        proofs.append(gc.sprove_begin(wires[rank], circ["public_inputs"], salts[rank], b, e, on_device=on_device))
    while True:
        qdist.exchange_local(proofs)
        if proofs[0].phase == 6:
            break
        for p in proofs:
            p.next()
    out = [p.end() for p in proofs]
    for c, ptrs in dev:
        for p in ptrs:
            c.dev_free(p)
    for gc in gcs:
        gc.free()
    for c in ctxs:
        c.close()
    return out


@pytest.mark.parametrize("k,zk,world", [(8, False, 2), (10, True, 2), (10, True, 4), (9, True, 8), (13, False, 8)])
def test_sharded_proof_equals_single_gpu_proof(k, zk, world):
    circ = minibuilder.build(k, zk=zk, seed=60 + k)
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=16)
    want = oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    got = _prove_sharded(circ, world, on_device=(world == 4))
    for rank, p in enumerate(got):
        assert p == want, "rank %d of %d" % (rank, world)
    rc, _ = orc.verify(circ["common"], oc.verifier_only_bytes(), got[0])
    assert rc == 0


def test_sharded_recursion_shaped_proof():
    """The aggregation-node gate set (14 gates) over 4 ranks: bytes equal to the oracle prover's."""
    class _Prov:
        poseidon_tables = staticmethod(orc.poseidon_tables)
        hash_no_pad = staticmethod(orc.hash_no_pad)

    circ = synth.build_recursion(10, zk=True, seed=11, provider=_Prov())
    oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=16)
    want = oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    for p in _prove_sharded(circ, 4):
        assert p == want


def test_full_range_is_the_plain_proof():
    """[0, 2^cap_height) on one rank: nothing to exchange, the phases are qpzk_prove in steps."""
    import qpzk
    circ = minibuilder.build(7, zk=True, seed=77)
    ctx = qpzk.Context(0)
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    want = gc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    sp = gc.sprove_begin(circ["wires"], circ["public_inputs"], circ["salts"], 0, 16)
    while sp.phase != 6:
        assert sp.exchanges() == []
        sp.next()
    assert sp.end() == want
    # unaligned shard: refused
    with pytest.raises(qpzk.QpzkError):
        gc.sprove_begin(circ["wires"], circ["public_inputs"], circ["salts"], 1, 2)
    # the handle is usable again
    assert gc.prove(circ["wires"], circ["public_inputs"], circ["salts"]) == want
    gc.free()
    ctx.close()

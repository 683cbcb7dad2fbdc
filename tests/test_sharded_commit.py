"""Multi-GPU sharded commit: cap subtrees = contiguous bit-reversed leaf ranges =
whole LDE cosets; the only exchange is the all-gather of the subtree roots.

CPU part (`-m "not gpu"`): shard arithmetic, and the world_size-2 all-gather over gloo with the oracle
standing in for the per-rank compute. GPU part: every rank's shard computed on cuda:0 through the C ABI
and assembled must equal the unsharded commit and the oracle, bit for bit."""
import os
import socket

import numpy as np
import pytest

from helpers import rand_felts
from oracle import oracle as orc
from qpzk import dist as qdist


def test_shard_ranges_cover_the_cap():
    for cap_h, r in ((4, 3), (4, 4), (2, 3), (0, 3), (5, 3)):
        units = (1 << cap_h) >> max(0, cap_h - r)
        for world in (1, 2, 4, 8):
            if world > units:
                with pytest.raises(ValueError):
                    qdist.shard_subtrees(0, world, cap_h, r)
                continue
            got = [qdist.shard_subtrees(rk, world, cap_h, r) for rk in range(world)]
            assert got[0][0] == 0 and got[-1][1] == 1 << cap_h
            for a, b in zip(got, got[1:]):
                assert a[1] == b[0]
            # whole cosets: a subtree range is a multiple of 2^(cap_h - r) subtrees
            unit = 1 << max(0, cap_h - r)
            assert all(b % unit == 0 and e % unit == 0 for b, e in got)
    with pytest.raises(ValueError):
        qdist.shard_subtrees(0, 3, 4, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(21)
        k, ncols, r, cap_h = 7, 5, 3, 4
        vals = rand_felts(rng, (ncols, 1 << k))
        full = orc.batch_commit(vals, r, cap_h, threads=1)
        b, e = qdist.shard_subtrees(rank, world, cap_h, r)
        # what a rank's shard holds: its own subtree roots, zeros elsewhere
        local = np.zeros_like(full["cap"])
        N = 1 << (k + r)
        per = N >> cap_h
        _, sub_cap = orc.merkle_new(full["leaves"][b * per:e * per], cap_h - (world.bit_length() - 1), threads=1)
        local[b:e] = sub_cap
        cap = qdist.allgather_cap_host(local, (b, e))
        q.put((rank, bool(np.array_equal(cap, full["cap"]))))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_cap_allgather():
    """world_size 2 on CPU: each rank builds only its half of the leaves' subtrees (oracle), the gloo
    all-gather assembles the cap every rank would observe into the transcript."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gloo_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(30)
    assert res == [(0, True), (1, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("k,ncols,salted", [(6, 5, False), (10, 12, True), (13, 135, False), (14, 20, True)])
def test_sharded_commit_matches_full(k, ncols, salted):
    import qpzk
    ctx = qpzk.Context(0)
    rng = np.random.default_rng(k * 100 + ncols)
    r, cap_h = 3, 4
    n, N = 1 << k, 1 << (k + r)
    vals = rand_felts(rng, (ncols, n))
    salts = rand_felts(rng, (4, N)) if salted else None
    want = orc.batch_commit(vals, r, cap_h, salts=salts, threads=8)
    d = ctx.dev_alloc(vals.nbytes)
    ctx.h2d(d, vals)
    ds = None
    if salted:
        p = ctx.dev_alloc(salts.nbytes)
        ctx.h2d(p, salts)
        ds = (p, 4)
    for world in (2, 4, 8):
        cap = np.zeros((1 << cap_h, 4), np.uint64)
        shards = []
        for rank in range(world):
            b, e = qdist.shard_subtrees(rank, world, cap_h, r)
            sh = qpzk.PolynomialBatch.from_values_shard_dev(ctx, d, ncols, n, r, cap_h, b, e, salts=ds)
            local = sh.cap
            assert not local[:b].any() and not local[e:].any()      # foreign subtrees are left zero
            cap[b:e] = local[b:e]
            shards.append((sh, b, e))
        assert np.array_equal(cap, want["cap"]), world
        # owned leaves open against the assembled cap; coefficients are complete on every rank
        per = N >> cap_h
        for sh, b, e in shards:
            sh.set_cap(cap)
            assert np.array_equal(sh.cap, want["cap"])
            for leaf in (b * per, e * per - 1, int(rng.integers(b * per, e * per))):
                row, sib = sh.open(leaf)
                assert np.array_equal(row, want["leaves"][leaf])
                assert orc.merkle_verify(row, leaf, cap, sib)
            assert np.array_equal(sh.polynomials, want["coeffs"])
            sh.free()
    # unaligned ranges are refused, not silently mis-computed
    with pytest.raises(qpzk.QpzkError):
        qpzk.PolynomialBatch.from_values_shard_dev(ctx, d, ncols, n, r, cap_h, 1, 2, salts=ds)
    ctx.dev_free(d)
    if ds:
        ctx.dev_free(ds[0])
    ctx.close()

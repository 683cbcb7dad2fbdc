"""The aggregator's tree schedule (/root/reference/wormhole/aggregator/src/circuits/tree.rs:55-103) dealt over
ranks: level structure, dependency order, and the world_size-2 gloo run in which each rank proves only its own
nodes and the per-level all-gather hands the children to whichever rank proves the parent. CPU only: the "node
proof" here is a hash of the children, so the result is checkable against a single-process run."""
import hashlib
import os
import socket

import pytest

from qpzk import aggregate as agg


def _node(level, index, children):
    h = hashlib.sha256(b"%d:%d" % (level, index))
    for c in children:
        h.update(c)
    return h.digest() * 4          # fixed-length "proof"


def _run(rank, world, gather, grouped=False):
    order = []

    def begin(level, index, children, slot):
        order.append((level, index, slot))
        return (level, index, list(children))

    def prove_group(level, index, children, ranks):
        # every rank of the group computes the same bytes, as the ranks of a sharded proof do
        order.append((level, index, tuple(ranks)))
        return _node(level, index, list(children))

    root, levels = agg.aggregate_tree([bytes([i]) * 128 for i in range(8)], 2, begin, lambda h: _node(*h), rank, world, gather,
                                      prove_group=prove_group if grouped else None)
    return root, levels, order


def test_levels_match_the_reference_defaults():
    assert agg.tree_levels(8, 2) == [4, 2, 1]           # tree.rs:17-20 defaults: 7 node proofs
    assert agg.tree_levels(16, 2) == [8, 4, 2, 1]
    assert agg.tree_levels(8, 8) == [1]                 # a flat 8-ary node
    assert agg.tree_levels(9, 4) == [3, 1]
    assert agg.tree_levels(1, 2) == []
    with pytest.raises(ValueError):
        agg.tree_levels(8, 1)


def test_single_rank_order_and_root():
    root, levels, order = _run(0, 1, None)
    assert [len(l) for l in levels] == [4, 2, 1]
    assert [o[0] for o in order] == [0, 0, 0, 0, 1, 1, 2]        # a level only after the one below it
    want1 = [_node(0, i, [bytes([2 * i]) * 128, bytes([2 * i + 1]) * 128]) for i in range(4)]
    assert levels[0] == want1
    assert root == _node(2, 0, [_node(1, 0, want1[:2]), _node(1, 1, want1[2:])])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, grouped=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if grouped:
            groups = agg.make_rank_groups(world, rank)
            assert sorted(groups) == [g for g in (2, 4, 8) if g <= world]
            assert all(dist.get_world_size(groups[g]) == g for g in groups)
        root, levels, order = _run(rank, world, agg.torch_all_gather(128), grouped)
        q.put((rank, root, [o if grouped else o[:2] for o in order]))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_tree():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(30)
    want_root, _, _ = _run(0, 1, None)
    assert res[0][1] == want_root and res[1][1] == want_root
    # nodes are dealt round-robin: rank 0 proves the even nodes of every level (and the root), rank 1 the odd ones
    assert res[0][2] == [(0, 0), (0, 2), (1, 0), (2, 0)]
    assert res[1][2] == [(0, 1), (0, 3), (1, 1)]


def test_ranks_per_node():
    # 8 leaves, branching factor 2 -> levels of 4, 2, 1 nodes: idle ranks join the upper levels
    assert [agg.ranks_per_node(c, 8) for c in (4, 2, 1)] == [2, 4, 8]
    assert [agg.ranks_per_node(c, 4) for c in (4, 2, 1)] == [1, 2, 4]
    assert [agg.ranks_per_node(c, 2) for c in (4, 2, 1)] == [1, 1, 2]
    assert [agg.ranks_per_node(c, 1) for c in (4, 2, 1)] == [1, 1, 1]
    assert agg.ranks_per_node(1, 16) == 8            # a proof shards by whole cosets: at most 8 ranks
    assert agg.ranks_per_node(3, 8) == 2 and list(agg.node_ranks(2, 2)) == [4, 5]
    assert agg.ranks_per_node(5, 8) == 1


def test_single_rank_grouped_is_the_plain_schedule():
    assert _run(0, 1, None, grouped=True)[0] == _run(0, 1, None)[0]


def test_gloo_world2_tree_with_group_proofs():
    """World 2: the levels of 4 and 2 nodes are dealt round-robin as before, the root is proved by BOTH ranks
    together (prove_group), and every rank ends with the root of the single-process run."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(rk, 2, port, q, True)) for rk in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(30)
    want_root, _, _ = _run(0, 1, None)
    assert res[0][1] == want_root and res[1][1] == want_root
    assert res[0][2] == [(0, 0, 0), (0, 2, 1), (1, 0, 0), (2, 0, (0, 1))]
    assert res[1][2] == [(0, 1, 0), (0, 3, 1), (1, 1, 0), (2, 0, (0, 1))]


def test_gloo_world4_tree_with_group_proofs():
    """World 4 (the shape of the 4- and 8-GPU runs: sub-groups of different sizes on different levels): the level
    of 4 nodes is one node per rank, the level of 2 nodes is proved by the rank pairs (0, 1) and (2, 3), the root by
    all four ranks, and every rank ends with the root of the single-process run."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(rk, 4, port, q, True)) for rk in range(4)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(30)
    want_root, _, _ = _run(0, 1, None)
    assert all(r[1] == want_root for r in res)
    for rk in range(4):
        assert res[rk][2] == [(0, rk, 0), (1, rk // 2, (2 * (rk // 2), 2 * (rk // 2) + 1)), (2, 0, (0, 1, 2, 3))]


# ---- the operator interface: buffer semantics of /root/reference/wormhole/tests/src/aggregator/aggregator_tests.rs ----
def _aggregator(**kw):
    return agg.WormholeProofAggregator(dummy_proof=b"\xdd" * 128, **kw)


def _aggregate(a):
    def begin(level, index, children, slot):
        return (level, index, list(children))

    return a.aggregate(begin, lambda h: _node(*h))


def test_tree_aggregation_config_defaults():
    c = agg.TreeAggregationConfig()                      # tree.rs:17-20, 31-52
    assert (c.tree_branching_factor, c.tree_depth, c.num_leaf_proofs) == (2, 3, 8)
    assert agg.TreeAggregationConfig(8, 1).num_leaf_proofs == 8 and agg.TreeAggregationConfig(4, 2).num_leaf_proofs == 16


def test_push_proof_to_buffer():                         # aggregator_tests.rs:10-22
    a = _aggregator()
    a.push_proof(b"\x01" * 128)
    assert len(a.proofs_buffer) == 1


def test_push_proof_to_full_buffer():                    # aggregator_tests.rs:24-43
    a = _aggregator()
    for _ in range(a.config.num_leaf_proofs):
        a.push_proof(b"\x01" * 128)
    with pytest.raises(agg.AggregatorError, match="proof buffer is full"):
        a.push_proof(b"\x01" * 128)
    assert len(a.proofs_buffer) == a.config.num_leaf_proofs


def test_aggregate_pads_with_dummy_proofs_and_takes_the_buffer():
    """aggregate_single_proof / aggregate_proofs_into_tree (aggregator_tests.rs:45-100): one pushed proof is padded
    to 8 leaves with the dummy proof and reduced by 4 + 2 + 1 node proofs; the buffer is gone afterwards."""
    a = _aggregator()
    a.push_proof(b"\x01" * 128)
    root, levels = _aggregate(a)
    assert [len(l) for l in levels] == [4, 2, 1]
    want = agg.aggregate_tree([b"\x01" * 128] + [b"\xdd" * 128] * 7, 2, lambda l, i, c, s: (l, i, list(c)),
                              lambda h: _node(*h))[0]
    assert root == want
    assert a.proofs_buffer is None
    with pytest.raises(agg.AggregatorError, match="no proofs to aggregate"):
        _aggregate(a)
    a.push_proof(b"\x02" * 128)                          # aggregator.rs:58-60: a push after `take` starts a new buffer
    assert a.proofs_buffer == [b"\x02" * 128]


def test_aggregate_full_buffer_needs_no_dummy_and_flat_config():
    a = agg.WormholeProofAggregator().with_config(agg.TreeAggregationConfig(8, 1))   # tree.rs:39-46: one flat 8-ary node
    for i in range(8):
        a.push_proof(bytes([i]) * 128)
    root, levels = _aggregate(a)
    assert [len(l) for l in levels] == [1] and root == _node(0, 0, [bytes([i]) * 128 for i in range(8)])
    b = agg.WormholeProofAggregator()                    # no dummy proof supplied and a short buffer
    b.push_proof(b"\x01" * 128)
    with pytest.raises(agg.AggregatorError, match="dummy proof"):
        _aggregate(b)


def test_pad_rejects_too_many_proofs():                  # util.rs:18-20
    with pytest.raises(agg.AggregatorError, match="more than the maximum"):
        agg.pad_with_dummy_proofs([b"x"] * 9, 8, b"d")

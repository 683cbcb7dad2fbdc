"""Multi-GPU plumbing for the sharded commit: one process per GPU under
`torch.distributed`; the only data-path exchange is the all-gather of the 2^cap_height subtree roots
(16 x 32 bytes at cap_height 4). Everything else - coefficients, LDE cosets, leaf hashing, subtrees -
is rank-local. Independent proofs (batch proving) need no exchange at all.

The reference has no counterpart (it is single-process rayon); the shard boundaries follow from the
bit-reversed leaf order of `PolynomialBatch` / `MerkleTree::new` in qp-plonky2 1.1.1: cap subtree s is
the leaf range [s*N/2^h, (s+1)*N/2^h), i.e. a set of whole LDE cosets.
"""
import numpy as np


def shard_subtrees(rank, world, cap_height, rate_bits):
    """Contiguous range of cap subtrees owned by `rank`, in whole cosets. Raises if `world` does not
    divide the number of shardable units."""
    ncap = 1 << cap_height
    unit = 1 << max(0, cap_height - rate_bits)      # subtrees per coset block
    nunits = ncap // unit
    if world > nunits or nunits % world:
        raise ValueError("world size %d does not divide the %d cosets of the LDE domain" % (world, nunits))
    per = nunits // world
    return rank * per * unit, (rank + 1) * per * unit


class _DevArray:
    """CUDA array interface over a raw device pointer (int64 view of u64 words)."""

    def __init__(self, ptr, nwords):
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


_EXT_STREAMS = {}


def _ext_stream(ctx):
    """torch view of the qpzk context's CUDA stream (cached per context)."""
    import torch
    key = (ctx.device, ctx.stream)
    if key not in _EXT_STREAMS:
        _EXT_STREAMS[key] = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", ctx.device))
    return _EXT_STREAMS[key]


def allgather_cap_nccl(batch, group=None, sync=True):
    """In-place NCCL all-gather of the subtree roots on the device cap of `batch` (every rank ends up
    with the full cap), on the context's own stream. Returns the torch view of the cap (int64 words)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nwords = 4 << batch.cap_height
    t = torch.as_tensor(_DevArray(batch.cap_dev, nwords), device=torch.device("cuda", batch.ctx.device))
    b, e = batch.subtrees
    per = nwords // world
    if (b * 4, e * 4) != (rank * per, (rank + 1) * per):
        raise ValueError("subtree range does not match the rank's slot in the all-gather")
    # on the context's own stream: ordered after the commit, no host synchronisation in between
    with torch.cuda.stream(_ext_stream(batch.ctx)):
        dist.all_gather_into_tensor(t, t[rank * per:(rank + 1) * per], group=group)
    if sync:
        batch.ctx.sync()
    return t


def allgather_cap_host(cap_local, subtrees, group=None):
    """Host-side variant (any backend, e.g. gloo): every rank contributes the rows of `cap_local` it
    owns; returns the assembled [2^h][4] cap."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    b, e = subtrees
    mine = torch.from_numpy(np.ascontiguousarray(cap_local[b:e]).view(np.int64).copy())
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return torch.cat(parts, 0).numpy().view(np.uint64)


# ---- one proof over several GPUs (include/qpzk.h qpzk_sprove_*, BASELINE configs[4]) ----
def _stream_ctx(ctx):
    """torch stream context over the qpzk context's own CUDA stream, so that a collective is ordered after the
    phase just enqueued and before the next one without any host synchronisation."""
    import torch
    return torch.cuda.stream(_ext_stream(ctx))


def exchange_nccl(proof, group=None):
    """Perform the exchanges due after the current phase of `proof` (a qpzk.ShardedProof) with NCCL, in place on
    the device buffers and on the context's stream: all-gather of subtree roots / quotient values, sum of the
    opened rows."""
    import torch
    import torch.distributed as dist

    ctx = proof.circuit.ctx
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", ctx.device)
    with _stream_ctx(ctx):
        for kind, ptr, words, b, e in proof.exchanges():
            t = torch.as_tensor(_DevArray(ptr, words), device=dev)
            if kind == proof.ALLGATHER:
                per = words // world
                if (b, e) != (rank * per, (rank + 1) * per):
                    raise ValueError("shard does not match the rank's slot in the all-gather")
                dist.all_gather_into_tensor(t, t[b:e], group=group)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def exchange_local(proofs):
    """The same exchanges between `proofs` that live in ONE process (one context each, any device): copies through
    the host. This is how the sharding logic is tested on a single GPU."""
    lists = [p.exchanges() for p in proofs]
    for items in zip(*lists):
        kind, words = items[0][0], items[0][2]
        bufs = []
        for p, (_, ptr, w, b, e) in zip(proofs, items):
            a = np.zeros(w, np.uint64)
            p.circuit.ctx.d2h(a, ptr)
            bufs.append(a)
        if kind == 1:
            full = np.zeros(words, np.uint64)
            for a, (_, _, _, b, e) in zip(bufs, items):
                full[b:e] = a[b:e]
        else:
            full = np.zeros(words, np.uint64)
            with np.errstate(over="ignore"):
                for a in bufs:
                    full = full + a
        for p, (_, ptr, _, _, _) in zip(proofs, items):
            p.circuit.ctx.h2d(ptr, full)


def prove_sharded_nccl(circuit, wires, public_inputs, salts, cap_height, rate_bits, on_device=False, group=None):
    """One proof over the ranks of `group`: every rank calls this with the same circuit and witness on its own
    context; returns the proof bytes (identical on every rank, equal to the single-GPU proof)."""
    import torch.distributed as dist

    b, e = shard_subtrees(dist.get_rank(group), dist.get_world_size(group), cap_height, rate_bits)
    proof = circuit.sprove_begin(wires, public_inputs, salts, b, e, on_device=on_device)
    while True:
        exchange_nccl(proof, group)
        if proof.phase == 6:
            break
        proof.next()
    return proof.end()

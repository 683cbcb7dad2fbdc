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


def allgather_cap_nccl(batch, group=None):
    """In-place NCCL all-gather of the subtree roots on the device cap of `batch` (every rank ends up
    with the full cap). Returns the torch view of the cap (int64 words)."""
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
    batch.ctx.sync()  # the commit ran on the context's stream; NCCL uses torch's
    dist.all_gather_into_tensor(t, t[rank * per:(rank + 1) * per], group=group)
    torch.cuda.current_stream().synchronize()
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

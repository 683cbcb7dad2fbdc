"""The aggregator's tree schedule over GPUs.

`aggregate_to_tree` (/root/reference/wormhole/aggregator/src/circuits/tree.rs:55-77) reduces the leaf proofs
level by level: every chunk of `branching_factor` proofs becomes one node proof (`aggregate_chunk`,
tree.rs:105-142: build the recursion circuit, `prove`), the chunks of a level are independent - the reference
fans them out with rayon `par_chunks` (tree.rs:92-103) - and a level needs the proofs of the level below. With the
default configuration (8 leaves, branching factor 2, tree.rs:17-20) that is 4 + 2 + 1 node proofs in three
dependent steps.

Here the node proofs of a level are dealt round-robin to the ranks (one process per GPU) and, inside a rank, to
its contexts, which prove concurrently through qpzk_prove_begin / qpzk_prove_end; after each level the node
proofs are all-gathered (`torch.distributed`, fixed-size byte tensors) so that whichever rank proves a parent
holds its children. Building a node's circuit and generating its witness from the child proofs is host work
outside this backend: the caller supplies it as `begin_node`.

The upper levels have fewer nodes than there are GPUs (2, then 1, on a box of 8). With `prove_group` the ranks that
would idle join in: a level of `count` nodes on `world` ranks gives every node `ranks_per_node(count, world)`
consecutive ranks, which prove it together as ONE proof sharded by cap subtrees (qpzk_sprove_*,
`qpzk.dist.prove_sharded_nccl` over the sub-group) - every rank of the group ends up with the bytes of the
single-GPU proof. A 2^13-row node takes 6.1 ms on one GPU and 4.9 ms on two.
"""
import numpy as np


def tree_levels(num_leaves, branching_factor):
    """Node proofs per level, bottom-up: [4, 2, 1] for 8 leaves and branching factor 2."""
    if num_leaves < 1 or branching_factor < 2:
        raise ValueError("need at least one leaf and a branching factor of at least 2")
    out, n = [], num_leaves
    while n > 1:
        n = -(-n // branching_factor)
        out.append(n)
    return out


def node_owner(index, world):
    return index % world


def ranks_per_node(count, world, max_ranks=8):
    """How many ranks prove ONE node of a level with `count` nodes: 1 while the level has at least as many nodes as
    half the ranks, else the largest power of two g with count * g <= world (at most `max_ranks`: a proof shards by
    whole LDE cosets, 8 at rate_bits 3)."""
    g = 1
    while count * g * 2 <= world and g * 2 <= max_ranks:
        g *= 2
    return g


def node_ranks(index, g):
    """The ranks that prove node `index` together when every node gets g ranks."""
    return range(index * g, (index + 1) * g)


def aggregate_tree(leaf_proofs, branching_factor, begin_node, end_node, rank=0, world=1, all_gather=None,
                   prove_group=None, max_ranks=8):
    """Run the schedule. `begin_node(level, index, children, slot)` enqueues the proof of one node on this rank's
    `slot`-th context and returns a handle; `end_node(handle)` waits for it and returns the proof bytes.
    `all_gather(list_of_bytes_or_None_per_node, owners)` fills in the nodes other ranks proved (None for world == 1);
    owners[i] is the rank whose copy of node i is taken.
    `prove_group(level, index, children, ranks)`, if given, is called by every rank in `ranks` (len(ranks) > 1) for
    a node of a level with fewer nodes than ranks and returns the proof bytes on each of them.
    Returns (root proof, [proof lists per level])."""
    proofs = list(leaf_proofs)
    levels = []
    for level, count in enumerate(tree_levels(len(proofs), branching_factor)):
        g = ranks_per_node(count, world, max_ranks) if prove_group is not None else 1
        out = [None] * count
        if g > 1:
            i = rank // g
            if i < count:
                children = proofs[i * branching_factor:(i + 1) * branching_factor]
                out[i] = prove_group(level, i, children, node_ranks(i, g))
            owners = [i * g for i in range(count)]
        else:
            mine = [i for i in range(count) if node_owner(i, world) == rank]
            handles = []
            for slot, i in enumerate(mine):
                children = proofs[i * branching_factor:(i + 1) * branching_factor]
                handles.append((i, begin_node(level, i, children, slot)))
            for i, h in handles:
                out[i] = end_node(h)
            owners = [node_owner(i, world) for i in range(count)]
        if world > 1:
            out = all_gather(out, owners)
        if any(p is None for p in out):
            raise RuntimeError("level %d is missing node proofs" % level)
        levels.append(out)
        proofs = out
    return proofs[0], levels


def make_rank_groups(world, rank, max_ranks=8):
    """torch.distributed sub-groups of g = 2, 4, .. consecutive ranks (every rank must call this: new_group is
    collective). Returns {g: the group this rank belongs to}."""
    import torch.distributed as dist
    groups, g = {}, 2
    while g <= world and g <= max_ranks:
        for j in range(world // g):
            ranks = list(range(j * g, (j + 1) * g))
            grp = dist.new_group(ranks)
            if rank in ranks:
                groups[g] = grp
        g *= 2
    return groups


def torch_all_gather(proof_len, device=None, group=None):
    """`all_gather` for aggregate_tree over torch.distributed: every level's node proofs have the same length
    (one circuit per level), so each rank contributes a [slots][proof_len] byte tensor holding, in node order, the
    nodes it is the owner of."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def gather(out, owners=None):
        count = len(out)
        if owners is None:
            owners = [node_owner(i, world) for i in range(count)]
        owned = [[i for i in range(count) if owners[i] == r] for r in range(world)]
        slots = max(1, max(len(o) for o in owned))
        mine = torch.zeros((slots, proof_len), dtype=torch.uint8)
        for s, i in enumerate(owned[rank]):
            if out[i] is None:
                raise ValueError("rank %d owns node %d but has no proof for it" % (rank, i))
            if len(out[i]) != proof_len:
                raise ValueError("node proofs of one level must have equal length")
            mine[s] = torch.frombuffer(bytearray(out[i]), dtype=torch.uint8)
        if device is not None:
            mine = mine.to(device)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        res = list(out)
        for r, t in enumerate(parts):
            t = t.cpu().numpy()
            for s, i in enumerate(owned[r]):
                res[i] = t[s].tobytes()
        return res

    return gather


def proof_digest(proof):
    """Short stand-in for "the parent's witness depends on this child" in tests of the schedule."""
    a = np.frombuffer(proof, dtype=np.uint8).astype(np.uint64)
    return int((a * (np.arange(a.size, dtype=np.uint64) + np.uint64(1))).sum() & np.uint64(0xFFFFFFFFFFFF))

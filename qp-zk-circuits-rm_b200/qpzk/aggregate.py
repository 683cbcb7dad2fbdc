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


def aggregate_tree(leaf_proofs, branching_factor, begin_node, end_node, rank=0, world=1, all_gather=None):
    """Run the schedule. `begin_node(level, index, children, slot)` enqueues the proof of one node on this rank's
    `slot`-th context and returns a handle; `end_node(handle)` waits for it and returns the proof bytes.
    `all_gather(list_of_bytes_or_None_per_node)` fills in the nodes other ranks proved (None for world == 1).
    Returns (root proof, [proof lists per level])."""
    proofs = list(leaf_proofs)
    levels = []
    for level, count in enumerate(tree_levels(len(proofs), branching_factor)):
        mine = [i for i in range(count) if node_owner(i, world) == rank]
        handles = []
        for slot, i in enumerate(mine):
            children = proofs[i * branching_factor:(i + 1) * branching_factor]
            handles.append((i, begin_node(level, i, children, slot)))
        out = [None] * count
        for i, h in handles:
            out[i] = end_node(h)
        if world > 1:
            out = all_gather(out)
        if any(p is None for p in out):
            raise RuntimeError("level %d is missing node proofs" % level)
        levels.append(out)
        proofs = out
    return proofs[0], levels


def torch_all_gather(proof_len, device=None, group=None):
    """`all_gather` for aggregate_tree over torch.distributed: every level's node proofs have the same length
    (one circuit per level), so each rank contributes a [slots][proof_len] byte tensor."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def gather(out):
        count = len(out)
        slots = -(-count // world)
        mine = torch.zeros((slots, proof_len), dtype=torch.uint8)
        for s in range(slots):
            i = s * world + rank
            if i < count and out[i] is not None:
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
            for s in range(slots):
                i = s * world + r
                if i < count:
                    res[i] = t[s].tobytes()
        return res

    return gather


def proof_digest(proof):
    """Short stand-in for "the parent's witness depends on this child" in tests of the schedule."""
    a = np.frombuffer(proof, dtype=np.uint8).astype(np.uint64)
    return int((a * (np.arange(a.size, dtype=np.uint64) + np.uint64(1))).sum() & np.uint64(0xFFFFFFFFFFFF))

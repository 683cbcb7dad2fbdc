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


# ---- the aggregator's operator interface (host side of the aggregation path) ----
DEFAULT_TREE_BRANCHING_FACTOR = 2   # tree.rs:17
DEFAULT_TREE_DEPTH = 3              # tree.rs:20


class AggregatorError(Exception):
    """What the reference reports with `anyhow::bail!` (aggregator.rs:53-55, 76-78; util.rs:18-20)."""


class TreeAggregationConfig:
    """`TreeAggregationConfig::new` (/root/reference/wormhole/aggregator/src/circuits/tree.rs:31-52):
    num_leaf_proofs = tree_branching_factor ^ tree_depth; the default is 2 ^ 3 = 8 leaves."""

    def __init__(self, tree_branching_factor=DEFAULT_TREE_BRANCHING_FACTOR, tree_depth=DEFAULT_TREE_DEPTH):
        if tree_branching_factor < 2 or tree_depth < 1:
            raise ValueError("need a branching factor of at least 2 and a depth of at least 1")
        self.tree_branching_factor = int(tree_branching_factor)
        self.tree_depth = int(tree_depth)
        self.num_leaf_proofs = self.tree_branching_factor ** self.tree_depth


def pad_with_dummy_proofs(proofs, proof_len, dummy_proof):
    """`pad_with_dummy_proofs` (/root/reference/wormhole/aggregator/src/util.rs:11-30): the buffer is filled up to
    `proof_len` proofs with copies of the dummy proof the reference ships (aggregator/data/dummy_proof[_zk].bin,
    passed in by the caller as bytes); more than `proof_len` proofs is an error."""
    proofs = list(proofs)
    if len(proofs) > proof_len:
        raise AggregatorError("proofs to aggregate was more than the maximum allowed")
    if len(proofs) < proof_len and dummy_proof is None:
        raise AggregatorError("failed to deserialize dummy proof")
    return proofs + [dummy_proof] * (proof_len - len(proofs))


class WormholeProofAggregator:
    """Host-side mirror of `WormholeProofAggregator` (/root/reference/wormhole/aggregator/src/aggregator.rs:13-93):
    the same buffer semantics and error behaviour, with `aggregate` running the tree schedule of this module over
    the ranks and contexts of the caller instead of `aggregate_to_tree`'s rayon fan-out. Proofs are serialized
    `ProofWithPublicInputs` bytes. Building a node circuit and its witness from the children stays with the caller
    (`begin_node` / `end_node` / `prove_group`, see `aggregate_tree`)."""

    def __init__(self, config=None, dummy_proof=None):
        self.config = config if config is not None else TreeAggregationConfig()
        self.dummy_proof = dummy_proof
        self.proofs_buffer = []          # Some(Vec::with_capacity(num_leaf_proofs)), aggregator.rs:30

    def with_config(self, config):
        self.config = config
        return self

    def push_proof(self, proof):
        """aggregator.rs:51-63: an error once the buffer holds num_leaf_proofs; after `aggregate` took the buffer a
        push starts a new one."""
        if self.proofs_buffer is not None:
            if len(self.proofs_buffer) >= self.config.num_leaf_proofs:
                raise AggregatorError("tried to add proof when proof buffer is full")
            self.proofs_buffer.append(proof)
        else:
            self.proofs_buffer = [proof]

    def aggregate(self, begin_node, end_node, rank=0, world=1, all_gather=None, prove_group=None, max_ranks=8):
        """aggregator.rs:74-92: takes the buffer, pads it with dummy proofs and reduces it to the root proof.
        Returns (root proof, [proof lists per level])."""
        if self.proofs_buffer is None:
            raise AggregatorError("there are no proofs to aggregate")
        proofs, self.proofs_buffer = self.proofs_buffer, None
        padded = pad_with_dummy_proofs(proofs, self.config.num_leaf_proofs, self.dummy_proof)
        return aggregate_tree(padded, self.config.tree_branching_factor, begin_node, end_node, rank, world, all_gather,
                              prove_group=prove_group, max_ranks=max_ranks)

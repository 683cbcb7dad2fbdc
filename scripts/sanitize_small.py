#!/usr/bin/env python3
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): every kernel family once, at
sizes where the instrumented run finishes in a minute - commits through both NTT paths (single-CTA and
two-pass), salted and unsalted, a sharded commit, a Merkle tree over row-major leaves, and one proof."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits-rm_b200"))
import numpy as np  # noqa: E402
import qpzk  # noqa: E402
from qpzk import synth  # noqa: E402

ctx = qpzk.Context(0)
rng = np.random.default_rng(0)
for k, ncols, salted in ((5, 3, False), (10, 9, True), (13, 5, False)):
    vals = rng.integers(0, qpzk.P, size=(ncols, 1 << k), dtype=np.uint64)
    salts = rng.integers(0, qpzk.P, size=(4, 1 << (k + 3)), dtype=np.uint64) if salted else None
    b = qpzk.PolynomialBatch.from_values(ctx, vals, 3, 4, salts=salts)
    b.cap, b.open(17), b.get_lde_values([0, 5, 9], 8)
    b.export()
    b.free()
d = ctx.dev_alloc(vals.nbytes)
ctx.h2d(d, vals)
sh = qpzk.PolynomialBatch.from_values_shard_dev(ctx, d, ncols, 1 << k, 3, 4, 4, 8)
sh.cap
sh.free()
ctx.dev_free(d)
t = qpzk.MerkleTree(ctx, rng.integers(0, qpzk.P, size=(256, 11), dtype=np.uint64), 2)
t.cap, t.prove(200), t.digests
t.free()
circ = synth.build(7, zk=True, seed=3, provider=synth.GpuProvider(ctx))
gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
proof = gc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
print("ok", len(proof))
gc.free()
ctx.close()

#!/usr/bin/env python3
"""k_tree_climb alone: MerkleTree::new over 2^k four-element leaves (hash_or_noop copies them, so the launch list
shows the climb from 2^(k-1) nodes to the cap by itself). usage: prof_climb.py [k] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits-rm_b200"))
import numpy as np  # noqa: E402
import qpzk  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 13
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = qpzk.Context(0)
rng = np.random.default_rng(0)
leaves = rng.integers(0, qpzk.P, size=(1 << k, 4), dtype=np.uint64)
for _ in range(reps):
    t0 = time.perf_counter()
    t = qpzk.MerkleTree(ctx, leaves, 4)
    print("2^%d leaves: %.1f us wall" % (k, (time.perf_counter() - t0) * 1e6))
    t.free()
ctx.close()

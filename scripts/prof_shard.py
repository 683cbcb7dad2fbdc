#!/usr/bin/env python3
"""One rank's part of a sharded 2^16 x 135 commit, alone on one GPU (usage: prof_shard.py world [reps]): the stage
times a rank of `world` GPUs sees, without NCCL - where does the fixed cost of the sharded commit sit?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits-rm_b200"))
import numpy as np  # noqa: E402
import qpzk  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
k, ncols = 16, 135
ctx = qpzk.Context(0)
rng = np.random.default_rng(0)
vals = rng.integers(0, qpzk.P, size=(ncols, 1 << k), dtype=np.uint64)
d = ctx.dev_alloc(vals.nbytes)
ctx.h2d(d, vals)
per = 16 // world
for _ in range(reps):
    t0 = time.perf_counter()
    b = qpzk.PolynomialBatch.from_values_shard_dev(ctx, d, ncols, 1 << k, 3, 4, 0, per)
    wall = (time.perf_counter() - t0) * 1e3
    print("world %d: wall %.3f ms, stages %s" % (world, wall, ctx.stage_ms()))
    b.free()
ctx.close()

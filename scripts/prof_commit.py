#!/usr/bin/env python3
"""Tiny driver for ncu: a few device-resident commits of a given shape (no oracle, no torch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits-rm_b200"))
import numpy as np  # noqa: E402
import qpzk  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
ncols = int(sys.argv[2]) if len(sys.argv) > 2 else 135
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = qpzk.Context(0)
rng = np.random.default_rng(0)
vals = rng.integers(0, qpzk.P, size=(ncols, 1 << k), dtype=np.uint64)
d = ctx.dev_alloc(vals.nbytes)
ctx.h2d(d, vals)
for _ in range(reps):
    b = qpzk.PolynomialBatch.from_values_dev(ctx, d, ncols, 1 << k, 3, 4)
    print(ctx.stage_ms())
    b.free()
ctx.close()

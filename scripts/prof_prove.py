#!/usr/bin/env python3
"""Driver for timing / ncu: a few proofs of a synthetic wormhole-shaped circuit (no oracle)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits-rm_b200"))
import numpy as np  # noqa: E402
import qpzk  # noqa: E402
from qpzk import synth  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
zk = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = qpzk.Context(0)
t0 = time.time()
circ = synth.build(k, zk=zk, seed=1, provider=synth.GpuProvider(ctx))
print("synth %.1fs" % (time.time() - t0))
t0 = time.time()
c = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
print("circuit_create %.2f ms" % ((time.time() - t0) * 1e3), ctx.stage_ms())
for i in range(reps):
    l0 = ctx.launch_count()
    t0 = time.time()
    proof = c.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    dt = (time.time() - t0) * 1e3
    st = c.stage_ms()
    print("prove %.2f ms wall, launches %d, stages sum %.2f: %s" % (dt, ctx.launch_count() - l0, sum(st.values()),
          {a: round(b, 3) for a, b in st.items()}), len(proof))

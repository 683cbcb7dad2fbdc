#!/usr/bin/env python3
"""Tiny driver: single-stream proof latency and stage times for a synthetic circuit of 2^k rows."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits-rm_b200"))
import numpy as np  # noqa: E402
import qpzk  # noqa: E402
from qpzk import synth  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
zk = bool(int(sys.argv[2])) if len(sys.argv) > 2 else (k == 14)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = qpzk.Context(0)
circ = synth.build(k, zk=zk, seed=1, provider=synth.GpuProvider(ctx))
gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
d = ctx.dev_alloc(circ["wires"].nbytes)
ctx.h2d(d, circ["wires"])
ds = None
if circ["salts"] is not None:
    ds = []
    for s in circ["salts"]:
        p = ctx.dev_alloc(s.nbytes)
        ctx.h2d(p, s)
        ds.append(p)
lat = []
for i in range(reps + 3):
    t0 = time.perf_counter()
    proof = gc.prove_dev(d, circ["public_inputs"], ds)
    if i >= 3:
        lat.append((time.perf_counter() - t0) * 1e3)
st = gc.stage_ms()
print("k=%d zk=%d coop_max=%s latency median %.3f ms min %.3f | %s" % (
    k, zk, os.environ.get("QPZK_COOP_MAX", "default"), float(np.median(lat)), float(np.min(lat)),
    " ".join("%s=%.3f" % (a, b) for a, b in st.items())))

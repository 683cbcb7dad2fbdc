#!/usr/bin/env python3
"""Tiny driver for ncu launch lists: a few proofs of one synthetic circuit (usage: prof_one_proof.py k [recursion] [reps])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits-rm_b200"))
import qpzk  # noqa: E402
from qpzk import synth  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
rec = len(sys.argv) > 2 and sys.argv[2] == "1"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = qpzk.Context(0)
circ = (synth.build_recursion if rec else synth.build)(k, zk=True, seed=k, provider=synth.GpuProvider(ctx))
gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
d = ctx.dev_alloc(circ["wires"].nbytes)
ctx.h2d(d, circ["wires"])
ds = []
for s in circ["salts"]:
    p = ctx.dev_alloc(s.nbytes)
    ctx.h2d(p, s)
    ds.append(p)
for _ in range(reps):
    proof = gc.prove_dev(d, circ["public_inputs"], ds)
    print(len(proof), gc.stage_ms())

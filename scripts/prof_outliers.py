# Where do the occasional slow first proofs of a process come from? Times prove_begin (enqueue) and prove_end (wait) separately
# and prints the per-stage device times of every proof.
import os, subprocess, sys
CHILD = r"""
import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/qp-zk-circuits-rm_b200"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, qpzk
from qpzk import synth
ctx = qpzk.Context(0)
k = int(os.environ.get("K", "17"))
circ = synth.build_recursion(k, zk=True, seed=10, provider=synth.GpuProvider(ctx))
gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
for i in range(7):
    t0 = time.perf_counter()
    gc.prove_begin(circ["wires"], circ["public_inputs"], circ["salts"])
    t1 = time.perf_counter()
    p = gc.prove_end()
    t2 = time.perf_counter()
    st = gc.stage_ms()
    print("proof %d: begin %.1f ms end %.1f ms total %.1f | device stages sum %.1f  %s" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, sum(st.values()), {a: round(b, 1) for a, b in st.items() if b > 1}), flush=True)
"""
for rep in range(int(os.environ.get("PROCS", "4"))):
    print("== process", rep, flush=True)
    subprocess.run([sys.executable, "-c", CHILD])

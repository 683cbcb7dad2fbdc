# Flat aggregation nodes from PAGEABLE host buffers: the staged multi-threaded upload (default) against the
# driver's own pageable path (QPZK_H2D_THREADS=1). One process per setting: the knob is read once.
import os, subprocess, sys
CHILD = r"""
import os
KS = [int(x) for x in os.environ.get("KS", "16,17,18").split(",")]
REPS = int(os.environ.get("REPS", "4"))
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/qp-zk-circuits-rm_b200"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, qpzk
from qpzk import synth
ctx = qpzk.Context(0)
for k in KS:
    circ = synth.build_recursion(k, zk=True, seed=10, provider=synth.GpuProvider(ctx))
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    ref = gc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    lat = []
    for _ in range(REPS):
        t0 = time.perf_counter(); p = gc.prove(circ["wires"], circ["public_inputs"], circ["salts"]); lat.append((time.perf_counter() - t0) * 1e3)
        assert p == ref
    print("k=%d latency all %s min %.1f median %.1f ms  commit_wires %.2f  proof sha %s" % (k, [round(x, 1) for x in lat], min(lat), float(np.median(lat)), gc.stage_ms()["commit_wires"], __import__("hashlib").sha256(ref).hexdigest()[:16]), flush=True)
    gc.free()
"""
for threads in os.environ.get("THREADS", "1,4,2,6").split(","):
    print("== QPZK_H2D_THREADS=" + threads, flush=True)
    subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, QPZK_H2D_THREADS=threads))

import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/qp-zk-circuits-rm_b200"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, qpzk
from qpzk import synth
from oracle import oracle as orc
ctx = qpzk.Context(0)
for k, rec in ((15, True), (16, True), (16, False), (17, False)):
    t = time.time()
    build = synth.build_recursion if rec else synth.build
    circ = build(k, zk=True, seed=k, provider=synth.GpuProvider(ctx))
    tb = time.time() - t
    gc = qpzk.Circuit(ctx, circ["common"], circ["digest"], circ["constants_sigmas"])
    proof = gc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    lat = []
    for _ in range(3):
        t0 = time.perf_counter(); gc.prove(circ["wires"], circ["public_inputs"], circ["salts"]); lat.append((time.perf_counter() - t0) * 1e3)
    rc, _ = orc.verify(circ["common"], gc.verifier_only_bytes(), proof)
    print("k=%d recursion=%s build %.1fs proof %d B verify rc=%d latency %.1f ms stages %s" % (k, rec, tb, len(proof), rc, min(lat), {a: round(b, 2) for a, b in gc.stage_ms().items()}), flush=True)
    gc.free()

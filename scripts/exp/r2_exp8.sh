#!/bin/bash
# cluster NTT for 2^13 / 2^14-point transforms (the wormhole proof sizes) against the single-CTA kernel
for m in 15 14 13; do
QPZK_NTT_CLUSTER_MIN=$m python - <<PY
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'qp-zk-circuits-rm_b200')
import numpy as np, qpzk, bench
from oracle import oracle as orc
ctx=qpzk.Context(0)
for k in (13,14):
    n=1<<k
    tr=bench.splitmix_trace(0x5EED0001+k,135,n); d=ctx.dev_alloc(tr.nbytes); ctx.h2d(d,tr)
    best=None
    for _ in range(6):
        b=qpzk.PolynomialBatch.from_values_dev(ctx,d,135,n,3,4); st=ctx.stage_ms(); cap=b.cap; b.free()
        if best is None or st['lde']<best['lde']: best=st
    want=orc.batch_commit(tr,3,4,threads=16)['cap'] if k==13 else None
    print('cluster_min=$m k=%d ifft %.3f lde %.3f leaf %.3f'%(k,best['ifft'],best['lde'],best['leaf_hash']), 'cap ok' if want is None or np.array_equal(cap,want) else 'CAP MISMATCH')
    ctx.dev_free(d)
PY
done

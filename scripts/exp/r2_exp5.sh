#!/bin/bash
# round 2, batch 5: cluster NTT v2 (radix-16 across the cluster, 3 CTAs per SM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
for on in 1 0; do
QPZK_NTT_CLUSTER=$on python - <<PY
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'qp-zk-circuits-rm_b200')
import numpy as np, qpzk, bench
ctx=qpzk.Context(0)
for k in (15,16):
    n=1<<k
    tr=bench.splitmix_trace(0x5EED0001+k,135,n); d=ctx.dev_alloc(tr.nbytes); ctx.h2d(d,tr)
    best=None
    for _ in range(6):
        b=qpzk.PolynomialBatch.from_values_dev(ctx,d,135,n,3,4); st=ctx.stage_ms(); b.free()
        if best is None or st['lde']<best['lde']: best=st
    print('cluster=$on k=%d ifft %.3f lde %.3f leaf %.3f levels %.3f'%(k,best['ifft'],best['lde'],best['leaf_hash'],best['merkle_levels']))
    ctx.dev_free(d)
PY
done
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:k_ntt_cluster -s 1 -c 1 -o gpurun_out/r2_ntt_cluster python scripts/prof_commit.py 16 135 2 > gpurun_out/r2_ntt_cluster.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_ntt_cluster.ncu-rep > gpurun_out/r2_ntt_cluster.txt 2>&1
python scripts/ncu_source_top.py gpurun_out/r2_ntt_cluster.ncu-rep 30 > gpurun_out/r2_ntt_cluster.top.txt 2>&1
cat gpurun_out/r2_ntt_cluster.txt; head -14 gpurun_out/r2_ntt_cluster.top.txt

#!/bin/bash
# round 2, batch 3: coop permutation with the duplicated exchange buffer; bench with 1 / 2 host threads
mkdir -p gpurun_out
for b in build/clat_*; do [ -x "$b" ] && timeout 120 $b $(basename $b); done > gpurun_out/r2e3.log 2>&1
cat gpurun_out/r2e3.log
timeout 600 python -m pytest tests/test_gpu_prover.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for ht in 1 2; do
timeout 600 python bench.py --no-aggregator --no-cpu --host-threads $ht > gpurun_out/r2e3_bench_ht$ht.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/r2e3_bench_ht$ht.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ht=$ht value',p['value'],'e2e',p['e2e']['value'],'lat',p['single_proof_latency_ms'],'launches',p['gpu_launches'], 'voting', p['voting_single_proof']['latency_ms_median'])
        print(p['proof_stage_ms']); print(p['commit_microbench']['stage_ms'])
PY
done

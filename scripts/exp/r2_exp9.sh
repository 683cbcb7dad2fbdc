#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2e9_tests.log
timeout 600 python bench.py --no-aggregator > gpurun_out/r2e9_bench.log 2>&1; tail -c 300 gpurun_out/r2e9_bench.log
python - <<PY
import json
for l in open('gpurun_out/r2e9_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value',p['value'],'e2e',p['e2e'],'lat',p['single_proof_latency_ms'])
        print(p['cpu_baseline']); print(p['commit_microbench']['stage_ms'], p['commit_microbench']['blinding']['stage_ms'])
        print({k:v for k,v in p['roofline_int'].items() if k.startswith('frac')}, p['roofline']['frac'], p['roofline']['traffic'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -c 900

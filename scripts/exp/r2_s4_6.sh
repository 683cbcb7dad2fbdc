#!/bin/bash
# fourth session of round 2, last call: loads requested four ahead in k_eval_at_ext / k_fri_compose, openings chunked from 2^12 coefficients
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_prover.py tests/test_gpu_checked_build.py tests/test_recursion_gates.py -m gpu -x -q 2>&1 | tail -2
timeout 60 python bench.py --steps 8 --warmup 4 --no-cpu --no-aggregator > gpurun_out/s4_6_short.log 2>/dev/null
python - <<PY
import json
for l in open('gpurun_out/s4_6_short.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value %.1f e2e %.1f lat %.3f voting %.3f' % (p['value'], p['e2e']['value'], p['single_proof_latency_ms'], p['voting_single_proof']['latency_ms_median']))
        print(p['proof_stage_ms'])
PY

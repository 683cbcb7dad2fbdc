#!/bin/bash
# fourth session of round 2, run 1: the 16-lane permutation with linearised partial rounds against the textbook form
# (bit-exact check + latency), the GPU suite, and a short bench
mkdir -p gpurun_out
SECONDS=0
timeout 120 build/coop_lat_lin0 textbook > gpurun_out/s4_1_coop_lat.log 2>&1
timeout 120 build/coop_lat_lin1 linear >> gpurun_out/s4_1_coop_lat.log 2>&1; cat gpurun_out/s4_1_coop_lat.log
echo "lat wall ${SECONDS}s"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_1_tests.log 2>&1; echo "tests rc=$? wall ${SECONDS}s"; tail -4 gpurun_out/s4_1_tests.log
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu --no-aggregator > gpurun_out/s4_1_short.log 2> gpurun_out/s4_1_short.err; echo "bench rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/s4_1_short.err
python - <<PY
import json
for l in open('gpurun_out/s4_1_short.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value',p['value'],'e2e',p['e2e']['value'],'lat',p['single_proof_latency_ms'])
        print('stages',p['proof_stage_ms'])
        print('voting',p['voting_single_proof']['latency_ms_median'], p['voting_single_proof']['stage_ms'])
PY

#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_7_tests.log 2>&1; echo "tests rc=$? wall=$SECONDS s"; tail -3 gpurun_out/s2_7_tests.log
timeout 1500 python bench.py > gpurun_out/s2_7_bench.log 2> gpurun_out/s2_7_bench.err; echo "bench rc=$? wall: $SECONDS s"; tail -3 gpurun_out/s2_7_bench.err
python - <<PY
import json
for l in open('gpurun_out/s2_7_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value',p['value'],'e2e',p['e2e']['value'],'lat',p['single_proof_latency_ms'])
        print('voting',p['voting_single_proof']['latency_ms_median'])
        print('commit', p['commit_microbench']['ms'], p['commit_microbench']['stage_ms'])
        a=p['aggregator_node_proof']
        print('node',a['latency_ms_median'], a['proofs_per_s_8_streams'], [ (k,v['latency_ms_median']) for k,v in a.items() if k.startswith('flat')])
        print('tree',p.get('aggregation_tree'))
        print('cpu',p['cpu_baseline'])
        print('roofline', p['roofline']['frac'], p['roofline_int']['frac'], p['roofline_int']['perms_per_s'])
PY

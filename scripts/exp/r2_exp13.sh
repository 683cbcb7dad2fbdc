#!/bin/bash
mkdir -p gpurun_out
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2e13_bench.log 2> gpurun_out/r2e13_bench.err; echo "bench wall: $SECONDS s"; tail -3 gpurun_out/r2e13_bench.err
python - <<PY
import json
for l in open('gpurun_out/r2e13_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value',p['value'],'e2e',p['e2e']['value'],'lat',p['single_proof_latency_ms'])
        print('voting',p['voting_single_proof']['latency_ms_median'])
        a=p['aggregator_node_proof']
        print('node',a['latency_ms_median'], [ (k,v['latency_ms_median'],v['stage_ms_median']) for k,v in a.items() if k.startswith('flat')])
        print('tree',p.get('aggregation_tree'))
        print('cpu',p['cpu_baseline'])
PY

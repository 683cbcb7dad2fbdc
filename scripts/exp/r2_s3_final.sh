#!/bin/bash
# third session of round 2: final N=1 line + the ncu launch list of the same command (short run), for profiles/
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python bench.py > gpurun_out/s3_bench.log 2> gpurun_out/s3_bench.err; echo "bench rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/s3_bench.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-aggregator > gpurun_out/s3_short.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aggregator > gpurun_out/s3_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/s3_launches.csv > gpurun_out/s3_launches_summary.txt; head -12 gpurun_out/s3_launches_summary.txt
python - <<PY
import json
for l in open('gpurun_out/s3_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value',p['value'],'e2e',p['e2e']['value'],'lat',p['single_proof_latency_ms'])
        print('voting',p['voting_single_proof']['latency_ms_median'])
        a=p['aggregator_node_proof']
        print('node',a['latency_ms_median'], [ (k,v['latency_ms_median']) for k,v in a.items() if k.startswith('flat')])
        print('micro',p['commit_microbench']['ms'], p['roofline']['frac'], p['roofline_int']['frac'])
        print('cpu',p['cpu_baseline'])
PY

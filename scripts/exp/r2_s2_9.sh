#!/bin/bash
# final N=1 line + the ncu launch list of the same command (short run), for profiles/
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/s2_9_bench.log 2> gpurun_out/s2_9_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s2_9_bench.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-aggregator > gpurun_out/s2_9_short.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/s2_9_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aggregator > gpurun_out/s2_9_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/s2_9_launches.csv > gpurun_out/s2_9_launches_summary.txt; head -30 gpurun_out/s2_9_launches_summary.txt
python - <<PY
import json
for l in open('gpurun_out/s2_9_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value',p['value'],'e2e',p['e2e']['value'],'lat',p['single_proof_latency_ms'])
        print('restored', p.get('circuit_restored_from_files'))
PY

#!/bin/bash
# fourth session of round 2: final check of HEAD on one B200 - GPU suite, smoke, the default bench line, the reference arm,
# the ncu launch list of a short bench, and ncu --set full of the linearised 16-lane permutation (transcript step, tree climb)
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests.log 2>&1; echo "tests rc=$? wall ${SECONDS}s"; tail -3 gpurun_out/s4_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4_smoke.log 2>&1; echo "smoke rc=$? wall ${SECONDS}s"; tail -1 gpurun_out/s4_smoke.log
timeout 1500 python bench.py > gpurun_out/s4_bench.log 2> gpurun_out/s4_bench.err; echo "bench rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/s4_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s4_ref.log 2> gpurun_out/s4_ref.err; echo "ref rc=$? wall ${SECONDS}s"; cut -c1-400 gpurun_out/s4_ref.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-aggregator > gpurun_out/s4_short.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/s4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aggregator > gpurun_out/s4_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/s4_launches.csv > gpurun_out/s4_launches_summary.txt; head -14 gpurun_out/s4_launches_summary.txt
echo "launch list wall ${SECONDS}s"
N="ncu --set full --clock-control none --import-source on -f"
cap() { # name, kernel regex, skip, command...
  name=$1; k=$2; s=$3; shift 3
  timeout 300 $N -k regex:$k -s $s -c 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt 2>&1
  python scripts/ncu_source_top.py gpurun_out/$name.ncu-rep 25 > gpurun_out/$name.top.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
cap s4_transcript_step k_transcript_step 5 python scripts/prof_one_proof.py 14 0 1
cap s4_tree_climb k_tree_climb 0 python scripts/prof_one_proof.py 9 0 1
echo "ncu full wall ${SECONDS}s"
python - <<PY
import json
for l in open('gpurun_out/s4_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value',p['value'],'e2e',p['e2e']['value'],'lat',p['single_proof_latency_ms'])
        print('voting',p['voting_single_proof']['latency_ms_median'])
        a=p['aggregator_node_proof']
        print('node',a['latency_ms_median'], a['proofs_per_s_8_streams'], [ (k,v['latency_ms_median']) for k,v in a.items() if k.startswith('flat')])
        print('tree',p['aggregation_tree']['latency_ms_median'])
        print('micro',p['commit_microbench']['ms'], p['roofline']['frac'], p['roofline_int']['frac'])
        print('cpu',p['cpu_baseline'])
PY

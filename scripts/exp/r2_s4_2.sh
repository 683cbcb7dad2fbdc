#!/bin/bash
# fourth session of round 2, run 2: table prefetch variants of the linearised 16-lane permutation, the climb thresholds
# revisited now that a 16-lane permutation is 24 % cheaper, chunk-ahead loads in the sponges
mkdir -p gpurun_out
SECONDS=0
for v in pf0_l10 pf1_l10 pf0_l11 pf1_l11; do timeout 120 build/coop_lat_$v $v 2>&1 | head -4; done > gpurun_out/s4_2_coop_lat.log; cat gpurun_out/s4_2_coop_lat.log
echo "lat wall ${SECONDS}s"
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu --no-aggregator > gpurun_out/s4_2_$name.log 2> gpurun_out/s4_2_$name.err
  python - "$name" <<PY
import json, sys
for l in open('gpurun_out/s4_2_%s.log' % sys.argv[1]):
    if l.startswith('{'):
        p=json.loads(l)
        print(sys.argv[1], 'value %.1f e2e %.1f lat %.3f voting %.3f' % (p['value'], p['e2e']['value'], p['single_proof_latency_ms'], p['voting_single_proof']['latency_ms_median']))
PY
}
run base
run pf QPZK_LIB=$PWD/build/libqpzk_pf.so
run coop2048 QPZK_COOP_MAX=2048
run coop4096 QPZK_COOP_MAX=4096
run leaf8192 QPZK_COOP_LEAF_MAX=8192
echo "wall ${SECONDS}s"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prover.py tests/test_gpu_checked_build.py -m gpu -x -q 2>&1 | tail -2
echo "wall ${SECONDS}s"

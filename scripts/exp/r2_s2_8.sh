#!/bin/bash
run() { echo "== $*"; env "$@" timeout 600 python bench.py --no-cpu --no-aggregator --steps ${STEPS:-16} --warmup 4 --streams ${STREAMS:-8} 2>/dev/null | tail -1 | python -c "import sys,json; p=json.loads(sys.stdin.read()); print('value', round(p['value'],1), 'e2e', round(p['e2e']['value'],1), 'lat', round(p['single_proof_latency_ms'],3), 'voting', round(p['voting_single_proof']['latency_ms_median'],3))"; }
run QPZK_COOP_MAX=256
STREAMS=12 run QPZK_COOP_MAX=1024
STREAMS=16 run QPZK_COOP_MAX=1024
STREAMS=6 run QPZK_COOP_MAX=1024
STREAMS=4 run QPZK_COOP_MAX=1024

#!/bin/bash
# QPZK_COOP_MAX sweep with the one-launch climb kernel: single-proof latency (2^14 ZK, 2^9) and throughput
mkdir -p gpurun_out
for cm in 512 1024 2048 4096; do
  echo "== QPZK_COOP_MAX=$cm"
  QPZK_COOP_MAX=$cm python scripts/prof_one_proof.py 14 0 4 2>&1 | tail -1 | python -c "import sys,ast; l=sys.stdin.read(); d=ast.literal_eval(l[l.index('{'):]); print('2^14 sum of stages %.3f ms' % sum(d.values()), {k: round(v,3) for k,v in d.items()})"
  QPZK_COOP_MAX=$cm python scripts/prof_one_proof.py 9 0 4 2>&1 | tail -1 | python -c "import sys,ast; l=sys.stdin.read(); d=ast.literal_eval(l[l.index('{'):]); print('2^9  sum of stages %.3f ms' % sum(d.values()), {k: round(v,3) for k,v in d.items()})"
  QPZK_COOP_MAX=$cm timeout 600 python bench.py --no-cpu --no-aggregator --steps 16 --warmup 4 2>/dev/null | tail -1 | python -c "import sys,json; p=json.loads(sys.stdin.read()); print('value', round(p['value'],1), 'e2e', round(p['e2e']['value'],1), 'lat', round(p['single_proof_latency_ms'],3), 'voting', round(p['voting_single_proof']['latency_ms_median'],3))"
done

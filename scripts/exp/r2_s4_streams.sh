#!/bin/bash
# fourth session of round 2, last GPU call: proofs in flight per GPU re-swept with the linearised 16-lane permutation
mkdir -p gpurun_out
for s in 6 8 10 12; do
  timeout 100 python bench.py --steps 8 --warmup 4 --streams $s --no-cpu --no-aggregator > gpurun_out/s4_streams_$s.log 2>/dev/null
  python - $s <<PY
import json, sys
for l in open('gpurun_out/s4_streams_%s.log' % sys.argv[1]):
    if l.startswith('{'):
        p=json.loads(l)
        print('streams', sys.argv[1], 'value %.1f e2e %.1f ms/step %.2f' % (p['value'], p['e2e']['value'], p['ms_per_step']))
PY
done

#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/s2_10_n$N.log 2> gpurun_out/s2_10_n$N.err
echo "rc=$?"; tail -3 gpurun_out/s2_10_n$N.err
python - <<PY
import json
for l in open('gpurun_out/s2_10_n$N.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value',p['value'],'e2e',p['e2e']['value'],'n',p['n_gpus'], 'clocks', p['clocks'])
        for k in ('aggregation_tree','commit_microbench_sharded','aggregation_proof_sharded'):
            print(k, json.dumps(p.get(k))[:1200])
PY

#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}; K=${2:-14}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${3:-20} --warmup 3 --shard-k $K --no-cpu > gpurun_out/r2e7_n$N.log 2>&1
tail -c 600 gpurun_out/r2e7_n$N.log; echo
python - <<PY
import json
for l in open('gpurun_out/r2e7_n$N.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('value',p['value'],'e2e',p['e2e']['value'],'n',p['n_gpus'])
        for k in ('aggregation_tree','commit_microbench_sharded','aggregation_proof_sharded'):
            print(k, json.dumps(p.get(k))[:900])
PY

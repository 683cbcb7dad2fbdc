for cfg in "spin 4" "spin 6" "yield 6" "blocking 6" "yield 8"; do
  set -- $cfg
  taskset -c 0-3 python bench.py --streams $2 --sync $1 --no-cpu --no-aggregator --steps 96 > gpurun_out/sync_$1_$2.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/sync_$1_$2.log"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("4 cores $1 x$2:", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["config"]["host_wait"], d["config"]["host_cores"], "lat", round(d["single_proof_latency_ms"],2))
PY
done
python bench.py --streams 6 --sync yield --no-cpu --no-aggregator --steps 96 > gpurun_out/sync_all_yield.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/sync_all_yield.log | head -2

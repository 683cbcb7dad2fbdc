#!/bin/bash
# round 2, batch 2: sync-free pipeline - GPU tests, smoke, bench
mkdir -p gpurun_out
for b in build/pexp_*; do [ -x "$b" ] && timeout 120 $b 17 $(basename $b); done > gpurun_out/r2e2.log 2>&1
cat gpurun_out/r2e2.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2e2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r2e2_smoke.log
timeout 600 python bench.py --no-aggregator --no-cpu > gpurun_out/r2e2_bench.log 2>&1; tail -c 1200 gpurun_out/r2e2_bench.log

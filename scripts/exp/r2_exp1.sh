#!/bin/bash
# round 2, experiment batch 1: Poseidon product spellings (throughput), dependent-permutation latency, GPU tests, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r2e1.log
for b in build/pexp_*; do [ -x "$b" ] && timeout 120 $b 17 $(basename $b); done >> gpurun_out/r2e1.log 2>&1
for b in build/clat_*; do [ -x "$b" ] && timeout 120 $b $(basename $b); done >> gpurun_out/r2e1.log 2>&1
cat gpurun_out/r2e1.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2e1_tests.log
timeout 600 python bench.py --no-aggregator --no-cpu > gpurun_out/r2e1_bench.log 2>&1; tail -c 1500 gpurun_out/r2e1_bench.log

#!/bin/bash
# session 2, run 2: 16-lane groups with full-warp / half-warp masks; the serialized-batch tests; voting-sized proof again
mkdir -p gpurun_out
timeout 120 build/coop_lat > gpurun_out/s2_2_coop_lat.log 2>&1; cat gpurun_out/s2_2_coop_lat.log
timeout 900 python -m pytest tests/test_serialized_batch.py tests/test_gpu_parity.py tests/test_gpu_prover.py tests/test_gpu_checked_build.py -m gpu -x -q > gpurun_out/s2_2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s2_2_tests.log
python scripts/prof_one_proof.py 9 0 3 2>&1 | tail -1
python scripts/prof_one_proof.py 14 0 3 2>&1 | tail -1

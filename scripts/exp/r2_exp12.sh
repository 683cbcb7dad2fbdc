#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_recursion_gates.py tests/test_gpu_prover.py tests/test_gpu_sharded_proof.py -m gpu -x -q 2>&1 | tail -5
python scripts/prof_one_proof.py 16 1 3 2>&1 | tail -2
python scripts/prof_one_proof.py 17 1 3 2>&1 | tail -1

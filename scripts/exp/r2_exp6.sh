#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded_proof.py -m gpu -x -q 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_sharded_proof.py 2>&1 | tail -5

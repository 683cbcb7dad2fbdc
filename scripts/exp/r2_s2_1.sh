#!/bin/bash
# session 2, run 1: full GPU suite + launch list of a voting-sized proof (is the 3 ms kernel time or launch gaps?)
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_1_tests.log 2>&1; echo "tests rc=$? wall=$SECONDS s"; tail -3 gpurun_out/s2_1_tests.log
python scripts/prof_one_proof.py 9 0 3 > gpurun_out/s2_1_prove9_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s2_1_launches_prove9.csv python scripts/prof_one_proof.py 9 0 3 > gpurun_out/s2_1_prove9_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/s2_1_launches_prove9.csv > gpurun_out/s2_1_launches_prove9_summary.txt 2>&1
tail -5 gpurun_out/s2_1_prove9_plain.log

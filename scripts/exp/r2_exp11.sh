#!/bin/bash
# launch lists (per-kernel device time) of one 2^14 wormhole proof and one 2^16 recursion-shaped proof
mkdir -p gpurun_out
python scripts/prof_one_proof.py 14 0 3 > gpurun_out/r2_prove14_plain.log 2>&1; tail -2 gpurun_out/r2_prove14_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_prove14.csv python scripts/prof_one_proof.py 14 0 3 > gpurun_out/r2_prove14_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_prove14.csv "python scripts/prof_one_proof.py 14 0 3 (circuit build + 3 proofs, 2^14 ZK wormhole shape)" | tee gpurun_out/r2_launches_prove14_summary.txt | head -40
python scripts/prof_one_proof.py 16 1 2 > gpurun_out/r2_prove16_plain.log 2>&1; tail -2 gpurun_out/r2_prove16_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_prove16.csv python scripts/prof_one_proof.py 16 1 2 > gpurun_out/r2_prove16_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_prove16.csv "python scripts/prof_one_proof.py 16 1 2 (circuit build + 2 proofs, 2^16 ZK recursion shape)" | tee gpurun_out/r2_launches_prove16_summary.txt | head -40

#!/bin/bash
mkdir -p gpurun_out
for w in 1 2 4 8; do python scripts/prof_shard.py $w 4 2>&1 | tail -2; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s2_3_launches_shard8.csv python scripts/prof_shard.py 8 3 > gpurun_out/s2_3_shard8_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/s2_3_launches_shard8.csv

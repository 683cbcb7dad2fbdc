#!/bin/bash
mkdir -p gpurun_out
for k in 13 12 10 8; do
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 40 --csv --log-file gpurun_out/s2_4_climb$k.csv python scripts/prof_climb.py $k 3 > /dev/null 2>&1
grep -E "k_tree_climb|k_merkle_level|k_leaf_hash" gpurun_out/s2_4_climb$k.csv | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | tail -4
done

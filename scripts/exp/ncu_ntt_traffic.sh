#!/bin/bash
# DRAM traffic of the four NTT launches of one 2^16 x 135 commit (summaries only)
N="ncu --set full --clock-control none -f"
for spec in "k_ntt_pass_a 3 prof_ntt_a_lde" "k_ntt_pass_b_transpose 1 prof_ntt_bt"; do
  set -- $spec
  $N -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 python scripts/prof_commit.py 16 135 2 > gpurun_out/$3.log 2>&1
  python scripts/ncu_summary.py gpurun_out/$3.ncu-rep > gpurun_out/$3.txt
  rm -f gpurun_out/$3.ncu-rep
  grep "kernel\|dram__bytes\|time_duration" gpurun_out/$3.txt
done

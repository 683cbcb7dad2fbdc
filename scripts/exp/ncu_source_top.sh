#!/bin/bash
# usage: ncu_source_top.sh <name> <kernel-regex> <skip> <command...>: top source lines by stall samples
name=$1; k=$2; s=$3; shift 3
ncu --set full --clock-control none --import-source on -f -k regex:$k -s $s -c 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
python scripts/ncu_source_top.py gpurun_out/$name.ncu-rep 60 > gpurun_out/$name.top.txt
python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt
rm -f gpurun_out/$name.ncu-rep
cat gpurun_out/$name.top.txt

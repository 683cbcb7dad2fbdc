#!/bin/bash
# round-1 closing measurements: default bench line, ncu launch list of the bench command, LDE pass A capture
python bench.py > gpurun_out/bench_default.log 2>&1
tail -c 400 gpurun_out/bench_default.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 8 --warmup 3 --streams 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:k_ntt_pass_a -s 3 -c 1 -o gpurun_out/prof_ntt_a_lde python scripts/prof_commit.py 16 135 2 > gpurun_out/prof_ntt_a_lde.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_ntt_a_lde.ncu-rep > gpurun_out/prof_ntt_a_lde.txt
python scripts/ncu_source_top.py gpurun_out/prof_ntt_a_lde.ncu-rep 30 > gpurun_out/prof_ntt_a_lde.top.txt
rm -f gpurun_out/prof_ntt_a_lde.ncu-rep
$N -k regex:k_ntt_pass_b_rows -s 1 -c 1 -o gpurun_out/prof_ntt_b python scripts/prof_commit.py 16 135 2 > gpurun_out/prof_ntt_b.log 2>&1
python scripts/ncu_source_top.py gpurun_out/prof_ntt_b.ncu-rep 30 > gpurun_out/prof_ntt_b.top.txt
rm -f gpurun_out/prof_ntt_b.ncu-rep
ls -la gpurun_out | head -30

#!/bin/bash
# Poseidon: occupancy variants; ncu of the base and the ALU-spelled s-box variants (which pipe is the limiter?)
mkdir -p gpurun_out
for b in build/pexp_*; do [ -x "$b" ] && timeout 120 $b 17 $(basename $b); done > gpurun_out/r2e10.log 2>&1
cat gpurun_out/r2e10.log
N="ncu --set full --clock-control none --import-source on -f"
for v in base sboxalu minb8; do
  $N -k regex:k_exp -s 2 -c 1 -o gpurun_out/r2_pexp_$v build/pexp_$v 17 > gpurun_out/r2_pexp_$v.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r2_pexp_$v.ncu-rep > gpurun_out/r2_pexp_$v.txt 2>&1
  ncu -i gpurun_out/r2_pexp_$v.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,r=rows[0],rows[1],rows[2]
for k in ('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','smsp__inst_executed.sum','sm__cycles_elapsed.max'):
    if k in h: print('  %-70s %s %s'%(k,r[h.index(k)],u[h.index(k)]))
" >> gpurun_out/r2_pexp_$v.txt
  rm -f gpurun_out/r2_pexp_$v.ncu-rep
  echo "== $v"; cat gpurun_out/r2_pexp_$v.txt
done

#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each), summarised on the box
# (scripts/ncu_summary.py); the .ncu-rep files are dropped except the leaf hash (64 MiB pull limit).
N="ncu --set full --clock-control none --import-source on -f"
cap() { # name, kernel regex, skip, command...
  name=$1; k=$2; s=$3; shift 3
  $N -k regex:$k -s $s -c 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY' >> gpurun_out/$name.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) > 2:
    h, u, r = rows[0], rows[1], rows[2]
    for k in ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
              "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.avg.per_cycle_active", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
              "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__maximum_warps_per_active_cycle_pct",
              "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "dram__bytes.sum"):
        if k in h:
            i = h.index(k); print("  %-70s %s %s" % (k, r[i], u[i]))
PY
  [ "$name" != "prof_leaf_v5" ] && rm -f gpurun_out/$name.ncu-rep
}
cap prof_leaf_v5 'k_leaf_hash$' 1 python scripts/prof_commit.py 16 135 2
cap prof_ntt_a_v2 k_ntt_pass_a 2 python scripts/prof_commit.py 16 135 2
cap prof_ntt_b_v2 k_ntt_pass_b_rows 1 python scripts/prof_commit.py 16 135 2
cap prof_ntt_small14 k_ntt_small 3 python scripts/prof_commit.py 14 135 2
cap prof_quotient_v2 k_quotient 1 python scripts/prof_prove.py 14 1 1
ls -la gpurun_out/

#!/bin/bash
# fourth session of round 2: ncu launch list of ONE 2^14 proof (circuit build + 2 proofs) with the linearised 16-lane
# permutation, and ncu --set full of the transcript step and a tree climb
mkdir -p gpurun_out
SECONDS=0
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/s4_one_proof_launches.csv python scripts/prof_one_proof.py 14 0 2 > gpurun_out/s4_one_proof_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/s4_one_proof_launches.csv > gpurun_out/s4_one_proof_launches_summary.txt; head -12 gpurun_out/s4_one_proof_launches_summary.txt
echo "launch list wall ${SECONDS}s"
N="ncu --set full --clock-control none --import-source on -f"
cap() { # name, kernel regex, skip, command...
  name=$1; k=$2; s=$3; shift 3
  timeout 120 $N -k regex:$k -s $s -c 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt 2>&1
  python scripts/ncu_source_top.py gpurun_out/$name.ncu-rep 25 > gpurun_out/$name.top.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
cap s4_transcript_step k_transcript_step 5 python scripts/prof_one_proof.py 14 0 1
echo "ncu full wall ${SECONDS}s"
cap s4_tree_climb k_tree_climb 0 python scripts/prof_one_proof.py 9 0 1
echo "ncu full wall ${SECONDS}s"
head -30 gpurun_out/s4_transcript_step.txt

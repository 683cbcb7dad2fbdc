#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient); usage: gpurun_retry.sh [gpurun args] -- cmd
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3

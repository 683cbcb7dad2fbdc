#!/bin/bash
# runs every build/pexp_* variant on the GPU box; results in gpurun_out/pexp.log
mkdir -p gpurun_out
for b in build/pexp_*; do [ -x "$b" ] && timeout 120 $b 17 $(basename $b); done 2>&1 | tee gpurun_out/pexp.log

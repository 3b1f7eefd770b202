#!/bin/bash
# last-panel width of the INT8-assisted factorisation; group-shape sweep of the variance kernel after the wide-N change
TAG=${1:-r2t}
mkdir -p gpurun_out
for LAST in 0 32 40 48 64; do
  GPR_FIT_LAST=$LAST timeout 300 python tools/fit_int8_prof.py 2>&1 | tail -1 | sed "s/^/last=$LAST /"
done
GRS="4 6 8 10 12 16 20" bash tools/oz_group_sweep.sh > /dev/null 2>&1
cat gpurun_out/oz_group_sweep.log

#!/bin/bash
# Final-state ncu evidence: launch list of the bench command + --set full of the INT8 kernels (variance, factorisation update,
# slicing) of the profiling target.  Plain runs first.
TAG=${1:-r2final}
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --fit-reps 1 --no-full-grid --no-fanout"
$BCMD > gpurun_out/plain_bench_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu list rc=$?"
python tools/prof_target.py > gpurun_out/plain_prof_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name regex:"ozaki_var_kernel|oz_slice_kernel|oz_diag_scale" -c 18 -f -o gpurun_out/prof_$TAG python tools/prof_target.py > gpurun_out/ncu_prof_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
rm -f gpurun_out/prof_$TAG.ncu-rep
ls -la gpurun_out/*$TAG* | head

#!/bin/bash
# Evidence run, part B (ncu): launch list of the bench command, --set full of the profiling target (fit + one variance batch).
# The .ncu-rep files are exported to CSV on the box and deleted (gpurun brings back at most 64 MiB).
TAG=${1:-r2c}
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --fit-reps 1 --no-full-grid --no-fanout"
$BCMD > gpurun_out/plain_bench_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu list rc=$?"
python tools/prof_target.py > gpurun_out/plain_prof_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -c 64 -f -o gpurun_out/prof_$TAG python tools/prof_target.py > gpurun_out/ncu_prof_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
for k in ozaki_var var_trsm chol_tiles; do
  ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:$k --print-source sass > gpurun_out/prof_${TAG}_src_$k.csv 2>/dev/null
done
rm -f gpurun_out/prof_$TAG.ncu-rep
ls -la gpurun_out | head -40; du -sh gpurun_out

#!/bin/bash
# GPU-box script: tests, bench (both arms), ncu launch list of the bench command, ncu --set full of one
# fit + one variance batch.  Outputs under gpurun_out/; summaries are copied to profiles/ by hand.
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$TAG.json
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --fit-reps 1"
$BCMD > gpurun_out/plain_bench_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu list rc=$?"
python tools/prof_target.py > gpurun_out/plain_prof_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -c 12 -f -o gpurun_out/prof_$TAG python tools/prof_target.py > gpurun_out/ncu_prof_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -5 gpurun_out/ncu_prof_$TAG.log

#!/bin/bash
# vectorised slicing kernel: exactness / parity tests, launch list of the bench command, short bench
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_parity.py -q -m gpu -x -k "int8 or cholesky or headline_parity_config3 or bit_identical or uninitialised or default_variance" 2>&1 | tail -4
BCMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fit-reps 2 --no-full-grid --no-fanout"
$BCMD > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --fit-reps 1 --no-full-grid --no-fanout > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu list rc=$?"

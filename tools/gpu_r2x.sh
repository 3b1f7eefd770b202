#!/bin/bash
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_parity.py -q -m gpu -x -k "int8 or cholesky or bit_identical or default_variance" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --fit-reps 1 --no-full-grid --no-fanout > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu list rc=$?"

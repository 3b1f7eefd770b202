#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$1.log
timeout 300 python tools/trace_chol.py 128 2>&1 | tail -8
timeout 300 python tools/quick_fit.py 4096 8192 16384
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench rc=$?"; cat gpurun_out/bench_$1.json

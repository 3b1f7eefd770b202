#!/bin/bash
# INT8-assisted fit as the default for n >= 8192: full GPU suite, config 5, the bench line.
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=6 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python tools/bench_configs.py 5 2 > gpurun_out/config5_$TAG.json 2> gpurun_out/config5_$TAG.err; echo "cfg5 rc=$?"; cut -c1-1800 gpurun_out/config5_$TAG.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench_$TAG.json

#!/bin/bash
# Evidence pass after the INT8-assisted fit and the wide-N MMA issue: full GPU suite, bench (20 steps, full), ncu.
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref_$TAG.json
bash tools/profile_ncu.sh $TAG
timeout 900 python tools/bench_configs.py 5 2 > gpurun_out/config5_$TAG.json 2> gpurun_out/config5_$TAG.err; echo "cfg5 rc=$?"; cut -c1-900 gpurun_out/config5_$TAG.json

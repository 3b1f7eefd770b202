#!/bin/bash
# Round 2, first GPU pass: the whole -m gpu suite (new: headline-size parity, forward-substitution variance, micro-batcher,
# fan-out bench), then a short bench run and the fan-out measurement with and without the micro-batcher.
TAG=${1:-r2a}
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 python tools/fanout_bench.py --unbatched --out gpurun_out/fanout_$TAG.json > /dev/null 2> gpurun_out/fanout_$TAG.err; echo "fanout rc=$?"; tail -4 gpurun_out/fanout_$TAG.err | cut -c1-1200

#!/bin/bash
# Final evidence pass of the round on one GPU: the whole -m gpu suite, the bench (both arms), then ncu (launch list + --set full).
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$TAG.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json
bash tools/profile_ncu.sh $TAG

#!/bin/bash
# Final evidence run of a round, part A (no profiler): tests, bench (both arms), per-config benches, accuracy report.
TAG=${1:-r1z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_$TAG.json
python tools/bench_append.py 8192 > gpurun_out/append_$TAG.json 2>&1; echo "append rc=$?"
python tools/bench_configs.py 1 2 > gpurun_out/configs12_$TAG.json 2> gpurun_out/configs12_$TAG.err; echo "cfg12 rc=$?"
python tools/bench_configs.py 5 2 > gpurun_out/config5_$TAG.json 2> gpurun_out/config5_$TAG.err; echo "cfg5 rc=$?"
python tools/accuracy_report.py 1024 2048 4096 > gpurun_out/accuracy_$TAG.json 2> gpurun_out/accuracy_$TAG.err; echo "acc rc=$?"

#!/bin/bash
# Last pass on one GPU (no profiler): smoke(), a 60-case stress sweep (a third of the fits forced INT8-assisted), the whole
# -m gpu suite, both bench arms.
TAG=${1:-r2final}
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log | cut -c1-300
timeout 900 python tools/stress.py 60 5000 > gpurun_out/stress_$TAG.log 2>&1; echo "stress rc=$?"; grep -c "fit=i8" gpurun_out/stress_$TAG.log; tail -3 gpurun_out/stress_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q --durations=4 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_$TAG.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_$TAG.json

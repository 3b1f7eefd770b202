#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$1.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$1.json; grep -o '"mean_panel_kernel_ms_per_step": [0-9.]*' gpurun_out/bench_$1.json
timeout 600 python tools/bench_configs.py 1 2 > gpurun_out/configs12_$1.json 2> gpurun_out/configs12_$1.err; echo "cfg12 rc=$?"; cat gpurun_out/configs12_$1.json; tail -3 gpurun_out/configs12_$1.err
timeout 900 python tools/bench_configs.py 5 2 > gpurun_out/config5_$1.json 2> gpurun_out/config5_$1.err; echo "cfg5 rc=$?"; cat gpurun_out/config5_$1.json; tail -3 gpurun_out/config5_$1.err

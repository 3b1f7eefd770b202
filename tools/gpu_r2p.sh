#!/bin/bash
# wide-N MMA issue (B slices concatenated along N): exactness tests, variance parity, fit study, bench.
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "int8 or cholesky" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "int8_tensor_cores or headline_parity_config3" 2>&1 | tail -5
timeout 900 python tools/fit_int8_study.py 16384 quick 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-full-grid --no-fanout > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_$TAG.json

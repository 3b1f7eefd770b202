#!/bin/bash
# k-chunked INT8 kernel: engine test, headline parity (config 3 and 5), config 5 bench.
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "int8" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "headline_parity or int8_tensor_cores or bit_identical" 2>&1 | tail -8
timeout 900 python tools/bench_configs.py 5 2 > gpurun_out/config5_$TAG.json 2> gpurun_out/config5_$TAG.err; echo "cfg5 rc=$?"; cut -c1-2500 gpurun_out/config5_$TAG.json
GPR_OZAKI_BASE=128 timeout 900 python tools/bench_configs.py 5 2 > gpurun_out/config5_base128_$TAG.json 2> gpurun_out/config5_base128_$TAG.err; echo "cfg5/128 rc=$?"; cut -c1-1200 gpurun_out/config5_base128_$TAG.json
cp gpurun_out/parity_full_size_n*.json gpurun_out/ 2>/dev/null

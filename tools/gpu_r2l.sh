#!/bin/bash
# INT8-assisted Cholesky: engine tests, then the panel-width / slice-count study at n = 16384.
TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "cholesky" 2>&1 | tail -5
timeout 900 python tools/fit_int8_study.py 16384 2>&1 | tail -12
cp gpurun_out/fit_int8_study_n16384.json gpurun_out/fit_int8_study_n16384_$TAG.json

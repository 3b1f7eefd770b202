#!/bin/bash
# INT8-assisted Cholesky: engine tests, then the panel-width / slice-count study at n = 16384.
TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "cholesky" -s 2>&1 | tail -25
timeout 900 python tools/fit_int8_study.py 16384 2>&1 | tail -12

#!/bin/bash
TAG=${1:-r2aa}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "cholesky" -s 2>&1 | grep -v "^$" | tail -12
timeout 600 python tools/fit_int8_prof.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "headline_parity_config3 or published or uninitialised" 2>&1 | tail -3

#!/bin/bash
# zero-slice skipping in the factorisation's INT8 updates: engine tests, fit timing (A/B by GPR_OZ_NOSKIP), parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "cholesky" 2>&1 | tail -2
timeout 300 python tools/fit_int8_prof.py 2>&1 | tail -2
GPR_OZ_NOSKIP=1 timeout 300 python tools/fit_int8_prof.py 2>&1 | tail -1 | sed 's/^/noskip: /'
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -q -m gpu -x -k "headline_parity_config3 or uninitialised or randomised" 2>&1 | tail -2

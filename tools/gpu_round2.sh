#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$1.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$1.log
timeout 120 python -c "
import gpr_b200 as g
for w in (0,1,2,3): print('peak', w, g.selftest_peak(w, 4))
"
timeout 300 python tools/trace_chol.py 128 2>&1 | tail -22
timeout 300 python tools/quick_fit.py 2048 4096 8192 16384
timeout 300 python tools/bench_append.py 8192 | tee gpurun_out/append_$1.json

#!/bin/bash
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -18 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$TAG.json'))
    for k in ('value','ms_per_step','e2e','roofline','variance_forms','time_to_first_variance_ms','time_until_variance_path_ready_ms','full_grid_s','fanout_calls_per_s','parity_full_size','fit_ms'):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print('no bench json', e)
PY
tail -5 gpurun_out/bench_$TAG.err

#!/bin/bash
# Round 2, two-GPU pass: the multi-device tests (in-process sharding, NCCL tail replica, factor published during the fit),
# the bench under torchrun at N=2, and the fan-out case again.
TAG=${1:-r2b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -3
timeout 900 python -m pytest tests -m gpu -x -q -k "two_devices or replicated or published or sample_on_chart or projection" > gpurun_out/pytest_multi_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_multi_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_$TAG.json 2> gpurun_out/bench_n2_$TAG.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n2_'+'$TAG'+'.json'))
    for k in ('value','ms_per_step','fit_ms','broadcast_ms','broadcast_GBps','broadcast_exposed_ms','fit_publish','full_grid','time_to_first_variance_ms'):
        print(k, d.get(k))
except Exception as e:
    print('no bench json', e)
PY
tail -5 gpurun_out/bench_n2_$TAG.err
timeout 300 python tools/fanout_bench.py --cases mugD:node,kettle:node --out gpurun_out/fanout_$TAG.json > /dev/null 2> gpurun_out/fanout_$TAG.err; echo "fanout rc=$?"; tail -2 gpurun_out/fanout_$TAG.err | cut -c1-900

#!/bin/bash
# 8-GPU validation: the publish check (factor replicated from inside the Cholesky kernel to 7 peers) and the bench under torchrun.
TAG=${1:-r2h}
N=${2:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/publish_check.py 6000 > gpurun_out/publish_n${N}_$TAG.log 2>&1; echo "publish rc=$?"; grep -E "publish:|PUBLISH_OK|Error|error" gpurun_out/publish_n${N}_$TAG.log | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${N}_$TAG.json'))
    for k in ('value','ms_per_step','e2e','fit_ms','broadcast_ms','broadcast_GBps','broadcast_exposed_ms','fit_publish','full_grid','clocks'):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print('no bench json', e)
PY
tail -5 gpurun_out/bench_n${N}_$TAG.err

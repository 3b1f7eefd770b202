#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shards or append" > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err

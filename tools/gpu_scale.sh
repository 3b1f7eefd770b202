#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shards" > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_multi_8.log
for N in 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; cut -c1-330 gpurun_out/bench_n$N.json; grep -o '"broadcast_ms": [0-9.]*' gpurun_out/bench_n$N.json
done

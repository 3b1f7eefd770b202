#!/bin/bash
# Round 2 evidence pass: ncu (launch list of the bench command + --set full of the profiling target), then config 5 and
# the 256^3 sampler on one GPU.
TAG=${1:-r2i}
mkdir -p gpurun_out
bash tools/profile_ncu.sh $TAG
timeout 900 python tools/bench_configs.py 5 2 > gpurun_out/config5_$TAG.json 2> gpurun_out/config5_$TAG.err; echo "cfg5 rc=$?"; cut -c1-1500 gpurun_out/config5_$TAG.json
timeout 600 python tools/bench_sampler.py > gpurun_out/sampler_$TAG.json 2> gpurun_out/sampler_$TAG.err; echo "sampler rc=$?"; cut -c1-800 gpurun_out/sampler_$TAG.json
timeout 600 python tools/bench_append.py 8192 > gpurun_out/append_$TAG.json 2>&1; echo "append rc=$?"; tail -c 600 gpurun_out/append_$TAG.json

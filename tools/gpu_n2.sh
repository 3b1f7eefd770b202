#!/bin/bash
# two-GPU pass: the multi-device tests and the bench under torchrun at N=2
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "two_devices or replicated or published" > gpurun_out/pytest_multi_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi_$TAG.log
bash tools/gpu_scale8.sh $TAG 2

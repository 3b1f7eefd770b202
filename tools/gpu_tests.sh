#!/bin/bash
TAG=${1:-t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_$TAG.log

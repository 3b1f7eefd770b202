#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tail_model or marching or chart or fanout or combined" > gpurun_out/pytest_new_$TAG.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_new_$TAG.log
bash tools/profile_ncu.sh $TAG

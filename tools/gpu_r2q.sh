#!/bin/bash
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 900 python tools/fit_int8_study.py 16384 2>&1 | tail -12
cp gpurun_out/fit_int8_study_n16384.json gpurun_out/fit_int8_study_n16384_$TAG.json

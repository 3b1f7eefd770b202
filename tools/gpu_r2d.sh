#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ozaki_check.py ${1:-3} > gpurun_out/ozaki_check.log 2>&1; echo "ozaki rc=$?"; tail -30 gpurun_out/ozaki_check.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader

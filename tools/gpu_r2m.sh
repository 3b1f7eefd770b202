#!/bin/bash
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 300 python tools/fit_int8_prof.py > gpurun_out/fit_int8_prof_plain_$TAG.log 2>&1; echo "plain rc=$?"; cat gpurun_out/fit_int8_prof_plain_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fit_int8_launches_$TAG.csv python tools/fit_int8_prof.py > gpurun_out/fit_int8_prof_ncu_$TAG.log 2>&1; echo "ncu rc=$?"

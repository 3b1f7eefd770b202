#!/bin/bash
# group shape of the variance kernel inside the real bench (power-capped steady state)
mkdir -p gpurun_out
for GR in 10 16 8 32 16 10; do
  GPR_OZ_GR=$GR timeout 600 python bench.py --steps 10 --warmup 3 --no-full-grid --no-fanout --no-cpu-baseline --fit-reps 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('gr=$GR value %.0f e2e %.0f ms/step %.2f clocks %s int8 TOP/s %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved']))"
done

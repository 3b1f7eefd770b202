#!/bin/bash
# Sweep of the co-scheduled group shape of ozaki_var_kernel (GPR_OZ_GR row tiles x 148/GR query tiles) at the headline size.
mkdir -p gpurun_out
for GR in ${GRS:-4 6 8 4 6 8 5 7}; do
  GPR_OZ_GR=$GR timeout 200 python - <<PY
import os, sys, json, numpy as np
sys.path.insert(0, os.getcwd())
import gpr_b200 as g
W = g.workloads
n = 16384
P, y, s2 = W.synthetic_cloud(n, seed=0)
ctx = g.Context(); reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
Q = W.grid_slab(256, 100, 104)[:4 * 148 * 128]
for rep in range(3):
    f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
t = ctx.timings()
print(json.dumps({"gr": $GR, "gq": 148 // $GR, "noskip": os.environ.get("GPR_OZ_NOSKIP", "0"), "int8_kernel_ms_per_batch": t["ozaki_ms"] / 4, "var_ms_per_batch": t["predict_var_ms"] / 4, "slices": t["ozaki_slices"], "issued_fraction": t["ozaki_issued_fraction"]}), flush=True)
PY
done 2>&1 | grep -v "^$" | tee gpurun_out/oz_group_sweep.log

"""Profiling target (GPU box): ONE fit at n=16384 (covariance build, tile Cholesky, two triangular solves),
the one-time L^-1, and ONE mean+variance batch of 148*128 queries through the host-pointer C-ABI call.
Run plain first, then under ncu (tools/profile_ncu.sh)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g

W = g.workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = g.Context()
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
P, y, s2 = W.synthetic_cloud(n, seed=0)
m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
print("fit", ctx.timings())
Q = W.grid_slab(256, 128, 129)[:148 * 128]
f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
print("predict", ctx.timings(), float(f.min()), float(v.min()), float(v.max()))

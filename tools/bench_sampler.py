"""The batched iso-surface sampler on the config-3 model (GPU box): the complete 256^3 lattice, mean everywhere, variance
for the |f| <= tol survivors only — against evaluating mean + variance for every lattice point (bench.py's rate)."""
import json
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g

W = g.workloads
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # devices of the one context: the sampler shards the lattice over them
ctx = g.Context(devices=list(range(ndev)))
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
P, y, s2 = W.synthetic_cloud(16384, seed=0)
m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
reg.prepare_variance(m)
h = W.SYNTH_GRID_HALF
step = 2 * h / (res - 1)
reg.sample_isosurface(m, lo=-0.1, hi=0.1, step=0.05, tol=0.01, var=False)      # warm-up: replicas on every device
if ndev > 1:
    reg.sample_isosurface(m, lo=-h, hi=h + 1e-9, step=2 * h / 127, tol=0.01)
t0 = time.perf_counter()
pts, f, v = reg.sample_isosurface(m, lo=-h, hi=h + 1e-9, step=step, tol=0.01)
wall = time.perf_counter() - t0
t = ctx.timings()
print(json.dumps({"workload": "n=16384 model, %d^3 lattice on [-%.1f,%.1f]^3, keep |f| <= 0.01 (node criterion), variance of the survivors" % (res, h, h),
                  "devices": ndev, "lattice_points": res ** 3, "kept": int(len(pts)), "wall_s": wall, "lattice_points_per_s": res ** 3 / wall,
                  "device_mean_ms_lattice": t["predict_mean_ms"], "device_var_ms_survivors": t["predict_var_ms"],
                  "var_min": float(v.min()), "var_max": float(v.max()), "abs_f_max": float(np.abs(f).max())}))

"""BASELINE config 4 (GPU box): n = 8192 base fit + 64 batches of 32 new surface points (label 0, sigma2 0.05,
src/gp_node.cpp:693) through the incremental append (gpr_append.cu) vs a full refit of the same points.
Prints one JSON line; copied to profiles/ by hand."""
import json
import os
import sys
import time

import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g

W = g.workloads
n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
batches, k = (int(sys.argv[2]) if len(sys.argv) > 2 else 64), 32
ctx = g.Context()
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
P, y, s2 = W.synthetic_cloud(n0, seed=0)
rng = np.random.default_rng(4)
New = rng.standard_normal((batches * k, 3))
New /= np.linalg.norm(New, axis=1, keepdims=True)          # on the unit sphere: inside the hull, R unchanged
yn, sn = np.zeros(len(New)), np.full(len(New), 0.05)

m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
fit_ms = ctx.timings()["fit_total_ms"]
reg.prepare_variance(m)
linv_ms = ctx.timings()["linv_ms"]
reg.reserve(m, n0 + batches * k)
dev, wall = [], []
for b in range(batches):
    s = slice(b * k, (b + 1) * k)
    t0 = time.perf_counter()
    reg.update(m, New[s, 0], New[s, 1], New[s, 2], yn[s], sn[s])
    wall.append(1e3 * (time.perf_counter() - t0))
    dev.append(ctx.timings()["append_ms"])
Pa = np.vstack([P, New]); ya = np.concatenate([y, yn]); sa = np.concatenate([s2, sn])
fresh = reg.create(Pa[:, 0], Pa[:, 1], Pa[:, 2], ya, sa)
refit_ms = ctx.timings()["fit_total_ms"]
reg.prepare_variance(fresh)
refit_linv_ms = ctx.timings()["linv_ms"]
a1, a2 = m.alpha, fresh.alpha
Q = W.grid_slab(24, 0, 24)[::7]
f1, v1 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
f2, v2 = reg.evaluate(fresh, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
n1 = n0 + batches * k
# algorithmic bytes of one append: the triangle of L^-1 (4 n^2 bytes) read twice for the new rows (B = X P, G = X^T B)
# and four times for alpha (X y and X^T v, for the solve and for its refinement step)
bytes_per = 6 * 4.0 * n1 * n1
print(json.dumps({"workload": "config4: n=%d base + %d x %d appended points, ThinPlate(R=%.1f)" % (n0, batches, k, W.SYNTH_R),
                  "base_fit_ms": fit_ms, "base_linv_ms": linv_ms,
                  "append_ms_device_median": float(np.median(dev)), "append_ms_device_first": dev[0], "append_ms_device_last": dev[-1],
                  "append_ms_wall_median": float(np.median(wall)),
                  "refit_ms_at_final_n": refit_ms, "refit_linv_ms_at_final_n": refit_linv_ms,
                  "speedup_vs_refit_fit_only": refit_ms / float(np.median(dev)),
                  "speedup_vs_refit_with_linv": (refit_ms + refit_linv_ms) / float(np.median(dev)),
                  "achieved_GBps_last": bytes_per / (dev[-1] * 1e-3) / 1e9,
                  "alpha_rel_vs_fresh_fit": rel(a1, a2), "mean_rel": rel(f1, f2), "var_rel": rel(v1, v2),
                  "final_n": n1}))

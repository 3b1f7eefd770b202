"""Second profiling target (GPU box): the latency-oriented paths — single-query fused kernel, incremental append,
indefinite-tail fit + predict, iso-surface sampler, projection — at n = 8192 (append / q = 1) and on mugD (tail)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g

W = g.workloads
ctx = g.Context()
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
P, y, s2 = W.synthetic_cloud(8192 + 32, seed=0)
m = reg.create(P[:8192, 0], P[:8192, 1], P[:8192, 2], y[:8192], s2[:8192])
reg.prepare_variance(m)
reg.reserve(m, 8192 + 128)
Q = W.grid_slab(64, 30, 31)
f1, v1 = reg.evaluate(m, Q[:1, 0], Q[:1, 1], Q[:1, 2], var=True)                                   # predict_small_kernel
reg.update(m, P[8192:, 0], P[8192:, 1], P[8192:, 2], y[8192:], s2[8192:])                          # append slab kernels + trsv
pts, fs, vs = reg.sample_isosurface(m, lo=-1.2, hi=1.2, step=2.4 / 63, tol=0.005)                 # grid fill / select + variance of survivors
out, st = reg.project(m, Q[:64] * 0.9, np.tile([0.0, 0.0, 1.0], (64, 1)), max_iter=50, step_mul=0.2)
cloud = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mugD_xyz.npy")).astype(np.float64)
Pn, yn, sn = W.node_training_set(cloud)
reg2 = g.GPRegressor("thin_plate", 2.0, ctx=ctx)
mn = reg2.create(Pn[:, 0], Pn[:, 1], Pn[:, 2], yn, sn)                                            # tail_* kernels
G = W.node_grid()
fn, vn = reg2.evaluate(mn, G[:, 0], G[:, 1], G[:, 2], var=True)
print("ok", float(v1[0]), len(pts), int((st > 0).sum()), mn.n_tail, float(vn.min()))

"""Small end-to-end workload that touches every kernel once; run under compute-sanitizer (one tool per gpurun call)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g

W = g.workloads
ctx = g.Context()
P, y, s2 = W.synthetic_cloud(700, seed=1)
rng = np.random.default_rng(1)
perm = rng.permutation(len(P)); P, y, s2 = P[perm], y[perm], s2[perm]
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
m = reg.create(P[:600, 0], P[:600, 1], P[:600, 2], y[:600], s2[:600], with_normals=True)          # cov, chol, trsv, normals
Q = W.grid_slab(24, 5, 7)                                                                        # 1152 queries
f, v, gr, tx, ty = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, tangent=True)           # linv, predict, var tiles, tangent
f1, v1, g1 = reg.evaluate(m, Q[:3, 0], Q[:3, 1], Q[:3, 2], var=True, grad=True)                  # fused small kernel
f0 = reg.evaluate(m, Q[:1, 0], Q[:1, 1], Q[:1, 2])
reg.update(m, P[600:650, 0], P[600:650, 1], P[600:650, 2], y[600:650], s2[600:650])            # incremental append (2 slabs, growth)
f2, v2 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
big = W.grid_slab(40, 0, 12)                                                                     # 19200 queries: thread kernel + split
fb, vb = reg.evaluate(m, big[:, 0], big[:, 1], big[:, 2], var=True)
pts, fs, vs = reg.sample_isosurface(m, lo=-1.2, hi=1.2, step=0.08, tol=0.01)
# indefinite tail (node setting)
cloud = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mugD_xyz.npy")).astype(np.float64)
Pn, yn, sn = W.node_training_set(cloud)
reg2 = g.GPRegressor("thin_plate", 2.0, ctx=ctx)
mn = reg2.create(Pn[:, 0], Pn[:, 1], Pn[:, 2], yn, sn)
fn, vn = reg2.evaluate(mn, Q[:, 0] * 0.8, Q[:, 1] * 0.8, Q[:, 2] * 0.8, var=True)
fq, vq = reg2.evaluate(mn, Q[:2, 0], Q[:2, 1], Q[:2, 2], var=True)
reg3 = g.GPRegressor("gaussian", 1.0, 1.0, ctx=ctx)
mg = reg3.create(P[:300, 0], P[:300, 1], P[:300, 2], y[:300], None)
fg, vg, gg = reg3.evaluate(mg, Q[:200, 0], Q[:200, 1], Q[:200, 2], var=True, grad=True)
for name, arr in (("f", f), ("v", v), ("grad", gr), ("tx", tx), ("f1", f1), ("v1", v1), ("f2", f2), ("v2", v2), ("fb", fb), ("vb", vb),
                  ("iso_f", fs), ("iso_v", vs), ("fn", fn), ("vn", vn), ("fq", fq), ("vq", vq), ("fg", fg), ("vg", vg), ("gg", gg)):
    assert np.isfinite(arr).all(), name
assert (v > 0).all() and (vb > 0).all() and (v2 > 0).all() and (vg > -1e-12).all()
print("ok", m.n, mn.n_tail, len(pts), float(np.abs(f).max()), float(v.min()), float(vn.min()), float(vg.min()))

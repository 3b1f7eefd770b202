import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import gpr_b200 as g
W = g.workloads
ctx = g.Context()
rng = np.random.default_rng(0)
# 1. Gaussian / Laplace kernels at scale, ragged n
for kind, p0, p1 in (("gaussian", 1.0, 1.0), ("laplace", 1.0, 0.7)):
    n = 20000
    P, y, s2 = W.synthetic_cloud(n, seed=1)
    reg = g.GPRegressor(kind, p0, p1, ctx=ctx)
    t0 = time.perf_counter(); m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2); wall = time.perf_counter() - t0
    t = ctx.timings()
    idx = np.arange(0, n, n // 50)
    d = np.sqrt(((P[idx, None, :] - P[None, :, :]) ** 2).sum(-1))
    K = (p0 * p0) * np.exp(-d / (p1 * p1)) if kind == "gaussian" else 2 * p0 * np.exp(-d / p1)
    K[np.arange(len(idx)), idx] += s2[idx]
    a = m.alpha
    Q = rng.uniform(-1.1, 1.1, (30000, 3))
    f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    print(kind, "n=%d fit %.1f ms (wall %.1f) chol %.1f ms resid %.2e | var range %.3g..%.3g finite=%s" % (
        n, t["fit_total_ms"], 1e3 * wall, t["chol_ms"], np.abs(K @ a - y[idx]).max(), v.min(), v.max(), np.isfinite(f).all() and np.isfinite(v).all()), flush=True)
    # append on a big ragged model
    Pn, yn, sn = W.synthetic_cloud(64, seed=5)
    reg.update(m, Pn[:40, 0] * 0.99, Pn[:40, 1] * 0.99, Pn[:40, 2] * 0.99, yn[:40], sn[:40])
    f2, v2 = reg.evaluate(m, Q[:500, 0], Q[:500, 1], Q[:500, 2], var=True)
    print("   after append n=%d append_ms=%.2f max|df|=%.2e var>0=%s" % (m.n, ctx.timings()["append_ms"], np.abs(f2 - f[:500]).max(), bool((v2 > 0).all())), flush=True)
    m.close()
# 2. very large mean-only / mean+grad batch through the host API
P, y, s2 = W.synthetic_cloud(4096, seed=2)
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
Q = rng.uniform(-1.2, 1.2, (5_000_000, 3))
t0 = time.perf_counter(); f, gr = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], grad=True); wall = time.perf_counter() - t0
f_small = reg.evaluate(m, Q[:1000, 0], Q[:1000, 1], Q[:1000, 2])
print("5M-query mean+grad call: %.2f s wall (%.1f M pts/s), matches small call: %.2e" % (wall, 5 / wall, np.abs(f[:1000] - f_small).max() / np.abs(f_small).max()))

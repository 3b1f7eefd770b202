"""Development tool (GPU box): fit timings at several n, with a host-side residual check."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g
W = g.workloads
ctx = g.Context()
for n in [int(a) for a in sys.argv[1:]] or [4096, 16384]:
    P, y, s2 = W.synthetic_cloud(n, seed=0)
    reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    best = None
    for rep in range(3):
        m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
        t = ctx.timings()
        if best is None or t["fit_total_ms"] < best["fit_total_ms"]:
            best = t
        a = m.alpha
        if rep < 2:
            m.close()
    idx = np.arange(0, n, max(1, n // 64))
    d = np.sqrt(((P[idx, None, :] - P[None, :, :]) ** 2).sum(-1))
    K = 2 * d ** 3 - 3 * W.SYNTH_R * d ** 2 + W.SYNTH_R ** 3
    K[np.arange(len(idx)), idx] += s2[idx]
    print(n, {k: round(best[k], 3) for k in ("cov_ms", "chol_ms", "solve_ms", "fit_total_ms")},
          "chol_tf=%.2f" % (n ** 3 / 3 / (best["chol_ms"] * 1e-3) / 1e12), "resid=%.2e" % np.abs(K @ a - y[idx]).max(), flush=True)
    m.close()

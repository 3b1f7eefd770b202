"""Accuracy attribution: GPU path and the double-precision CPU oracle, each against the
long-double oracle, on the synthetic config-3 cloud at several n; cond(K) beside each.  One JSON line per n."""
import json
import os
import sys
import time

import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g
import oracle

W = g.workloads
ctx = g.Context()
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
for n in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]:
    P, y, s2 = W.synthetic_cloud(n, seed=0)
    Q = W.grid_slab(32, 10, 12)[::3]
    t0 = time.time()
    ol = oracle.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", W.SYNTH_R, 0.0, factor="llt", precision="longdouble")
    fl, vl, gl = ol.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True, threads=os.cpu_count())
    od = oracle.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", W.SYNTH_R, 0.0, factor="llt")
    fd, vd, gd = od.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True, threads=os.cpu_count())
    oe = oracle.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", W.SYNTH_R, 0.0, factor="ldlt", dist="expansion")
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    fg, vg, gg = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    K = ol.get(K=True)["K"]
    w = np.linalg.eigvalsh(K)
    big = np.abs(fl) > 1e-9
    print(json.dumps({"n": n, "queries": len(Q), "cond_K": float(w.max() / w.min()), "cond_times_eps": float(w.max() / w.min() * 2.2e-16),
                      "gpu_vs_longdouble": {"alpha": rel(m.alpha, ol.alpha), "mean": rel(fg, fl), "var": rel(vg, vl), "grad": rel(gg, gl),
                                            "sign_mismatches": int((np.sign(fg) != np.sign(fl))[big].sum())},
                      "cpu_double_llt_vs_longdouble": {"alpha": rel(od.alpha, ol.alpha), "mean": rel(fd, fl), "var": rel(vd, vl), "grad": rel(gd, gl)},
                      "cpu_reference_arithmetic_ldlt_expansion_vs_longdouble": {"alpha": rel(oe.alpha, ol.alpha)},
                      "seconds": round(time.time() - t0, 1)}), flush=True)

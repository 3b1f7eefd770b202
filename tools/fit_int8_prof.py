"""One all-FP64 fit and two INT8-assisted fits at n = 16384 (target of an ncu launch list: per-launch times of the panel
factorisations, the INT8 updates and the slicing)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g
W = g.workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P, y, s2 = W.synthetic_cloud(n, seed=0)
ctx = g.Context()
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
for mode in ("fp64", "int8", "int8"):
    os.environ["GPR_FIT_MODE"] = mode
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    t = ctx.timings()
    print(mode, "chol %.2f ms, fit %.2f ms" % (t["chol_ms"], t["fit_total_ms"]), flush=True)
    m.close()

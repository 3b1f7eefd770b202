"""Where the first variance call of a process spends its time (GPU box): wall clock of each stage, twice."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g
W = g.workloads
n = 16384
P, y, s2 = W.synthetic_cloud(n, seed=0)
Q = W.grid_slab(256, 128, 129)[:148 * 128]
t0 = time.perf_counter(); ctx = g.Context(); print("ctx %.1f ms" % (1e3 * (time.perf_counter() - t0)))
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
for rep in range(3):
    t0 = time.perf_counter(); m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2); t1 = time.perf_counter()
    reg.prepare_variance(m); t2 = time.perf_counter()
    f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True); t3 = time.perf_counter()
    tt = ctx.timings()
    f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True); t4 = time.perf_counter()
    print("rep %d: fit %.1f | L^-1 %.1f (device %.1f) | first evaluate %.1f (var %.1f, int8 kernel %.1f) | second evaluate %.1f ms"
          % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), tt["linv_ms"], 1e3 * (t3 - t2), tt["predict_var_ms"], tt["ozaki_ms"], 1e3 * (t4 - t3)), flush=True)
    m.close()

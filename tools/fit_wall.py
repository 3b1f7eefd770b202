import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import gpr_b200 as g
W = g.workloads
ctx = g.Context(); reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
for n in (712, 4096, 16384):
    P, y, s2 = W.synthetic_cloud(n, seed=0)
    x, yy, z = np.ascontiguousarray(P[:, 0]), np.ascontiguousarray(P[:, 1]), np.ascontiguousarray(P[:, 2])
    m = None; rows = []
    for rep in range(6):
        if m is not None:
            t0 = time.perf_counter(); m.close(); tc = 1e3 * (time.perf_counter() - t0)
        else: tc = 0.0
        t0 = time.perf_counter(); m = reg.create(x, yy, z, y, s2); w = 1e3 * (time.perf_counter() - t0)
        rows.append((round(w, 2), round(ctx.timings()["fit_total_ms"], 2), round(tc, 2)))
    print(n, "wall/device/close-before ms:", rows, flush=True)
    m.close()

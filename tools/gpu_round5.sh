#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$1.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_$1.log
timeout 600 python tools/bench_configs.py 1 > gpurun_out/config1_$1.json 2> gpurun_out/config1_$1.err; echo "cfg1 rc=$?"; cat gpurun_out/config1_$1.json; tail -3 gpurun_out/config1_$1.err
timeout 300 python - <<'PY'
import time, numpy as np, gpr_b200 as g
W = g.workloads
ctx = g.Context(); reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
for n in (2048, 16384):
    P, y, s2 = W.synthetic_cloud(n, seed=0)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2); reg.prepare_variance(m)
    Q = W.grid_slab(16, 3, 4)
    for mode in ("mean", "mean+var"):
        for i in range(20): reg.evaluate(m, Q[i:i+1, 0], Q[i:i+1, 1], Q[i:i+1, 2], var=(mode != "mean"))
        t0 = time.perf_counter()
        for i in range(200): reg.evaluate(m, Q[i:i+1, 0], Q[i:i+1, 1], Q[i:i+1, 2], var=(mode != "mean"))
        print("n=%d q=1 %s: %.1f us per call (python ctypes wall), device %.1f us" % (n, mode, 1e6 * (time.perf_counter() - t0) / 200, 1e3 * (ctx.timings()["predict_var_ms"] + ctx.timings()["predict_mean_ms"])))
PY

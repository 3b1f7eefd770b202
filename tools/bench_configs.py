"""BASELINE.json configs 1, 2 and 5 on the GPU box (config 3 is bench.py, config 4 is tools/bench_append.py).
Prints one JSON line per config; the lines are copied to profiles/ by hand.

  python tools/bench_configs.py 1 2          # PCD clouds (tests/golden/*_xyz.npy), fit + predict, checked against the oracle
  python tools/bench_configs.py 5 [slabs]    # n = 65,536 fit on one GPU + a slab of the 512^3 variance sweep
"""
import json
import os
import sys
import time

import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpr_b200 as g

W = g.workloads
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def cloud(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + "_xyz.npy")).astype(np.float64)


def config1(ctx):
    """mugD.pcd, node preprocessing, ThinPlate(R = max pairwise distance), 29^3 node grid, mean+variance."""
    import oracle
    P, y, s2 = W.node_training_set(cloud("mugD"))
    R = W.max_pairwise_distance(P)
    reg = g.GPRegressor("thin_plate", R, ctx=ctx)
    Q = W.node_grid()
    best = None
    for _ in range(5):
        t0 = time.perf_counter()
        m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
        fit_wall = time.perf_counter() - t0
        tf = ctx.timings()
        t0 = time.perf_counter()
        f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        ev_wall = time.perf_counter() - t0
        tp = ctx.timings()
        cur = (fit_wall + ev_wall, fit_wall, ev_wall, tf["fit_total_ms"], tp["predict_total_ms"], tp["linv_ms"])
        best = cur if best is None or cur < best else best
    t0 = time.perf_counter()
    o = oracle.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", R, 0.0, factor="ldlt")
    cpu_fit = time.perf_counter() - t0
    t0 = time.perf_counter()
    fo, vo, _ = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, threads=os.cpu_count())
    cpu_ev = time.perf_counter() - t0
    # the node's own pattern: one query per call (src/gp_node.cpp:1074)
    t0 = time.perf_counter()
    for i in range(200):
        reg.evaluate(m, Q[i:i + 1, 0], Q[i:i + 1, 1], Q[i:i + 1, 2], var=True)
    single_us = 1e6 * (time.perf_counter() - t0) / 200
    ref_single_us = None
    if oracle.have_reference():      # the reference's own evaluate(), same call pattern, one host thread
        ref = oracle.Reference("thin_plate", R, 0.0).fit(P[:, 0], P[:, 1], P[:, 2], y, s2)
        t0 = time.perf_counter()
        for i in range(200):
            ref.evaluate(Q[i:i + 1, 0], Q[i:i + 1, 1], Q[i:i + 1, 2], 2)
        ref_single_us = 1e6 * (time.perf_counter() - t0) / 200
    return {"config": 1, "workload": "mugD.pcd (262 pts) + 15 external, ThinPlate(R=%.4f = max pairwise distance), 29^3 grid mean+var" % R,
            "n": len(P), "queries": len(Q), "gpu_fit_ms_device": best[3], "gpu_fit_ms_wall": 1e3 * best[1],
            "gpu_predict_ms_device": best[4], "gpu_predict_ms_wall": 1e3 * best[2], "gpu_queries_per_s_wall": len(Q) / best[2],
            "gpu_single_query_call_us": single_us, "reference_cpu_single_query_call_us": ref_single_us,
            "cpu_oracle_ldlt_fit_ms": 1e3 * cpu_fit, "cpu_oracle_predict_ms_%d_threads" % os.cpu_count(): 1e3 * cpu_ev,
            "alpha_rel": rel(m.alpha, o.alpha), "mean_rel": rel(f, fo), "var_rel": rel(v, vo),
            "sign_mismatches": int((np.sign(f) != np.sign(fo))[np.abs(fo) > 1e-9].sum())}


def config2(ctx):
    """kettle.pcd and jug.pcd, Gaussian(1,1), mean / variance / normal at all training points."""
    import oracle
    out = []
    for name in ("kettle", "jug"):
        P, y, s2 = W.node_training_set(cloud(name))
        reg = g.GPRegressor("gaussian", 1.0, 1.0, ctx=ctx)
        best = None
        for _ in range(5):
            t0 = time.perf_counter()
            m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2, with_normals=True)
            fit_wall = time.perf_counter() - t0
            tf = ctx.timings()
            t0 = time.perf_counter()
            f, v, gr = reg.evaluate(m, P[:, 0], P[:, 1], P[:, 2], var=True, grad=True)
            ev_wall = time.perf_counter() - t0
            tp = ctx.timings()
            cur = (fit_wall + ev_wall, fit_wall, ev_wall, tf["fit_total_ms"], tp["predict_total_ms"])
            best = cur if best is None or cur < best else best
        t0 = time.perf_counter()
        o = oracle.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "gaussian", 1.0, 1.0, factor="ldlt", with_normals=True)
        cpu_fit = time.perf_counter() - t0
        t0 = time.perf_counter()
        fo, vo, go = o.predict(P[:, 0], P[:, 1], P[:, 2], var=True, grad=True, threads=1)
        cpu_ev = time.perf_counter() - t0
        out.append({"cloud": name, "n": len(P), "gpu_fit_ms_device": best[3], "gpu_fit_ms_wall": 1e3 * best[1],
                    "gpu_predict_ms_device": best[4], "gpu_predict_ms_wall": 1e3 * best[2],
                    "cpu_oracle_ldlt_fit_ms": 1e3 * cpu_fit, "cpu_oracle_predict_ms_1_thread": 1e3 * cpu_ev,
                    "alpha_rel": rel(m.alpha, o.alpha), "mean_rel": rel(f, fo), "var_rel": rel(v, vo), "grad_rel": rel(gr, go),
                    "normals_abs": float(np.abs(m.get()["normals"] - o.get()["normals"]).max())})
    return {"config": 2, "workload": "kettle.pcd / jug.pcd + 15 external, Gaussian(1,1), sigma2=0.1, mean/var/normal at all training points",
            "clouds": out}


def config5(ctx, slabs=2):
    """n = 65,536 (K = 32 GiB in place), fit on one GPU, then `slabs` variance batches (148*128 queries each) of
    the 512^3 grid; the full-grid time is an explicit extrapolation."""
    n = 65536
    P, y, s2 = W.synthetic_cloud(n, seed=0)
    reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    t0 = time.perf_counter()
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    fit_wall = time.perf_counter() - t0
    tf = ctx.timings()
    reg.prepare_variance(m)
    linv_ms = ctx.timings()["linv_ms"]
    batch = 148 * 128
    Q = W.grid_slab(512, 256, 257)[:slabs * batch]
    f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)            # warm-up (allocates the 10 GB panel)
    t0 = time.perf_counter()
    f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    ev_wall = time.perf_counter() - t0
    tp = ctx.timings()
    # the two FP64 forms on the same batches (the default above is the INT8 tensor-core form)
    forms = {}
    for mode in ("product", "trsm"):
        os.environ["GPR_VAR_MODE"] = mode
        reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        tt = ctx.timings()
        forms[mode] = {"predict_var_ms": tt["predict_var_ms"], "fp64_tflops": float(n) ** 2 * len(Q) / (tt["predict_var_ms"] * 1e-3) / 1e12}
    os.environ.pop("GPR_VAR_MODE", None)
    # residual of the solve on a sample of rows (a CPU factorisation of 32 GiB is out of reach in a bench)
    idx = np.arange(0, n, n // 64)
    d = np.sqrt(((P[idx, None, :] - P[None, :, :]) ** 2).sum(-1))
    K = 2 * d ** 3 - 3 * W.SYNTH_R * d ** 2 + W.SYNTH_R ** 3
    K[np.arange(len(idx)), idx] += s2[idx]
    alpha = m.alpha
    resid = float(np.abs(K @ alpha - y[idx]).max())
    qps = len(Q) / (1e-3 * (tp["predict_mean_ms"] + tp["predict_var_ms"]))
    return {"config": 5, "workload": "synthetic n=65536, ThinPlate(R=4.2), fit on one GPU + %d variance batches of the 512^3 grid" % slabs,
            "n": n, "fit_ms_device": tf["fit_total_ms"], "fit_cov_ms": tf["cov_ms"], "fit_chol_ms": tf["chol_ms"], "fit_solve_ms": tf["solve_ms"],
            "fit_ms_wall": 1e3 * fit_wall, "chol_tflops": n ** 3 / 3 / (tf["chol_ms"] * 1e-3) / 1e12,
            "cov_GBps": 8.0 * n * (n + 128) / 2 / (tf["cov_ms"] * 1e-3) / 1e9,
            "linv_ms": linv_ms, "linv_tflops": n ** 3 / 3 / (linv_ms * 1e-3) / 1e12,
            "queries": len(Q), "predict_mean_ms": tp["predict_mean_ms"], "predict_var_ms": tp["predict_var_ms"],
            "var_fp64_equivalent_tflops": float(n) ** 2 * len(Q) / (tp["predict_var_ms"] * 1e-3) / 1e12,
            "int8_kernel_ms": tp["ozaki_ms"], "int8_slices": tp["ozaki_slices"], "fp64_forms": forms,
            "queries_per_s_device": qps, "queries_per_s_wall": len(Q) / ev_wall,
            "extrapolated_512^3_sweep_hours_1gpu": 512 ** 3 / qps / 3600, "extrapolated_512^3_sweep_minutes_8gpu": 512 ** 3 / qps / 8 / 60,
            "solve_residual_max_abs_on_%d_rows" % len(idx): resid, "var_min": float(v.min()), "var_max": float(v.max())}


if __name__ == "__main__":
    ctx = g.Context()
    args = sys.argv[1:] or ["1", "2"]
    if args[0] == "5":
        print(json.dumps(config5(ctx, *[int(a) for a in args[1:2]])), flush=True)
    else:
        for w in args:
            print(json.dumps({"1": config1, "2": config2}[w](ctx)), flush=True)

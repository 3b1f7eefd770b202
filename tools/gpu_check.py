"""Development check run on the GPU box: exercises every kernel against numpy / the oracle with verbose
diagnostics, without stopping at the first failure.  Not part of the product or of the test-suite."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpr_b200 as g          # noqa: E402
import oracle                 # noqa: E402

W = g.workloads
RES = {}


def step(name):
    def deco(fn):
        t0 = time.time()
        try:
            out = fn()
            RES[name] = {"ok": True, "out": out, "s": round(time.time() - t0, 2)}
        except Exception as e:     # noqa: BLE001
            RES[name] = {"ok": False, "err": repr(e), "tb": traceback.format_exc()[-1500:]}
        print(name, json.dumps(RES[name], default=str)[:1200], flush=True)
        return fn
    return deco


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@step("peaks")
def _():
    out = {}
    for c in (1, 2, 4, 8):
        out["dmma_tf_%d" % c] = round(g.selftest_peak(0, c), 2)
        out["dfma_tf_%d" % c] = round(g.selftest_peak(1, c), 2)
    return out


@step("gemm_mmajor")
def _():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((256, 64)); B = rng.standard_normal((384, 64))
    C = g.selftest_gemm(A, B, False)
    return rel(C, A @ B.T)


@step("gemm_kmajor")
def _():
    rng = np.random.default_rng(2)
    A = rng.standard_normal((128, 160)); B = rng.standard_normal((256, 160))
    C = g.selftest_gemm(A, B, True)
    return rel(C, A @ B.T)


def spd(n, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    return M @ M.T / n + np.eye(n)


@step("leaf")
def _():
    A = spd(128, 3)
    L, inv, info = g.selftest_leaf(A)
    Lr = np.linalg.cholesky(A)
    return {"info": info, "L": rel(np.tril(L), Lr), "upper0": float(np.abs(np.triu(L, 1)).max()),
            "inv": rel(inv, np.linalg.inv(Lr))}


@step("leaf_notspd")
def _():
    A = spd(128, 4); A[77, 77] = -5.0
    _, _, info = g.selftest_leaf(A)
    return info


for nbt, serial in ((1, True), (2, True), (3, True), (3, False), (8, False), (16, False)):
    @step("factor_nb%d_%s" % (nbt, "serial" if serial else "mega"))
    def _(nbt=nbt, serial=serial):
        A = spd(128 * nbt, 10 + nbt)
        L, X, piv = g.selftest_factor(A, True, serial)
        Lr = np.linalg.cholesky(A)
        return {"piv": piv, "L": rel(L, Lr), "Linv": rel(np.tril(X), np.linalg.inv(Lr))}


def fit_case(name, P, y, s2, kind, p0, p1, Q, normals=False):
    @step("fit_" + name)
    def _():
        reg = g.GPRegressor(kind, p0, p1)
        t0 = time.time()
        m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2, with_normals=normals)
        t_fit = time.time() - t0
        o = oracle.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, kind, p0, p1, factor="llt", dist="diff", with_normals=normals)
        out = {"n": len(P), "fit_s": round(t_fit, 3), "tim": reg.ctx.timings(), "alpha": rel(m.alpha, o.alpha), "R": m.R - o.R}
        og = o.get(K=False, factor=True)
        out["L"] = rel(m.factor(), np.tril(og["factor"]))
        if normals:
            out["normals"] = float(np.abs(m.get()["normals"] - og["normals"]).max())
        f, v, gr, tx, ty = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True, tangent=True)
        fo, vo, go = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True, threads=8)
        N, Tx, Ty = oracle.tangent_basis(go)
        out.update(f=rel(f, fo), v=rel(v, vo), vabs=float(np.abs(v - vo).max()), g=rel(gr, go), tx=float(np.abs(tx - Tx).max()),
                   ty=float(np.abs(ty - Ty).max()), sign=int((np.sign(f) != np.sign(fo)).sum()))
        f1 = reg.evaluate(m, Q[:5, 0], Q[:5, 1], Q[:5, 2])
        f2, v2 = reg.evaluate(m, Q[:3, 0], Q[:3, 1], Q[:3, 2], var=True)
        out.update(f_small=rel(f1, fo[:5]), v_small=rel(v2, vo[:3]))
        out["tim_pred"] = reg.ctx.timings()
        return out


mug = np.load(os.path.join(ROOT, "tests/golden/mugD_xyz.npy")).astype(np.float64)
P, y, s2 = W.node_training_set(mug)
Rm = W.max_pairwise_distance(P)
grid = W.node_grid()[::37]
fit_case("mugD_thinplate", P, y, s2, "thin_plate", Rm, 0.0, grid, normals=True)
ket = np.load(os.path.join(ROOT, "tests/golden/kettle_xyz.npy")).astype(np.float64)
P2, y2, s22 = W.node_training_set(ket)
fit_case("kettle_gaussian", P2, y2, s22, "gaussian", 1.0, 1.0, P2[::3], normals=True)
fit_case("kettle_laplace", P2, y2, s22, "laplace", 1.0, 1.0, P2[::5])


@step("notspd_R2")
def _():
    reg = g.GPRegressor("thin_plate", 2.0)
    try:
        reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    except g.GPRegressionException as e:
        return {"code": e.code, "pivot": e.pivot, "msg": str(e)[:120]}
    return "no error!"


P3, y3, s23 = W.synthetic_cloud(2048, seed=0)
fit_case("synth2048", P3, y3, s23, "thin_plate", W.SYNTH_R, 0.0, W.grid_slab(16, 0, 16)[::7])


@step("big_fit")
def _():
    out = {}
    for n in (4096, 8192, 16384):
        Pn, yn, sn = W.synthetic_cloud(n, seed=0)
        reg = g.GPRegressor("thin_plate", W.SYNTH_R)
        for rep in range(2):
            m = reg.create(Pn[:, 0], Pn[:, 1], Pn[:, 2], yn, sn)
            t = reg.ctx.timings()
            if rep == 0:
                m.close()
        a = m.alpha
        # residual check in float64 on the host for a row subset: (K alpha)_i = y_i
        idx = np.arange(0, n, max(1, n // 64))
        d = np.sqrt(((Pn[idx, None, :] - Pn[None, :, :]) ** 2).sum(-1))
        K = 2 * d ** 3 - 3 * W.SYNTH_R * d ** 2 + W.SYNTH_R ** 3
        K[np.arange(len(idx)), idx] += sn[idx]
        res = float(np.abs(K @ a - yn[idx]).max())
        out[n] = {"cov_ms": t["cov_ms"], "chol_ms": t["chol_ms"], "solve_ms": t["solve_ms"], "fit_ms": t["fit_total_ms"],
                  "chol_tf": n ** 3 / 3 / (t["chol_ms"] * 1e-3) / 1e12, "resid": res}
        Q = W.grid_slab(64, 0, 10)[:148 * 128]
        t0 = time.time(); reg.prepare_variance(m); out[n]["linv_ms"] = reg.ctx.timings()["linv_ms"]
        for rep in range(2):
            f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        tp = reg.ctx.timings()
        out[n].update(var_ms=tp["predict_var_ms"], mean_ms=tp["predict_mean_ms"], q=len(Q),
                      var_tf=n * n * len(Q) / (tp["predict_var_ms"] * 1e-3) / 1e12, vmin=float(v.min()), vmax=float(v.max()))
        fm = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2])
        out[n]["mean_only_ms"] = reg.ctx.timings()["predict_mean_ms"]
        out[n]["f_consistent"] = rel(fm, f)
        m.close()
    return out


try:
    import torch
    @step("cublas_dgemm")
    def _():
        out = {}
        for n in (4096, 8192):
            a = torch.randn(n, n, device="cuda", dtype=torch.float64); b = torch.randn(n, n, device="cuda", dtype=torch.float64)
            for _ in range(2):
                c = a @ b
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(3):
                c = a @ b
            e1.record(); torch.cuda.synchronize()
            out[n] = round(3 * 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12, 2)
        n = 16384
        a = torch.randn(n, n, device="cuda", dtype=torch.float64)
        torch.linalg.cholesky(a @ a.T / n + torch.eye(n, device="cuda", dtype=torch.float64))
        spdm = a @ a.T / n + torch.eye(n, device="cuda", dtype=torch.float64)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); torch.linalg.cholesky(spdm); e1.record(); torch.cuda.synchronize()
        out["cusolver_potrf_16384_ms"] = round(e0.elapsed_time(e1), 2)
        return out
except ImportError:
    pass

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as fh:
    json.dump(RES, fh, indent=1, default=str)
print("FAILED:", [k for k, v in RES.items() if not v["ok"]])

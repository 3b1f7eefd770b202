"""Randomised parity stress (GPU box): random sizes, kernels, noise settings, query-batch sizes, append sequences and
indefinite tails against the CPU oracle.  Prints one line per case and a summary; exit code 1 on any violation.
  python tools/stress.py [cases] [seed]"""
import os
import sys
import time
import traceback

import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g
import oracle

W = g.workloads
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = g.Context()
bad = []
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def cloud(rng, n):
    d = rng.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.where(rng.random(n) < 0.75, 1.0, 2.0) * rng.uniform(0.97, 1.03, n)
    P = d * r[:, None] * rng.uniform(0.6, 1.0, 3)
    y = np.where(r > 1.5, 1.0, 0.0) + 0.01 * rng.standard_normal(n)
    return P, y


t_start = time.time()
for case in range(cases):
    seed = seed0 + case
    rng = np.random.default_rng(seed)
    try:
        n = int(rng.choice([1, 2, 3, 17, 127, 128, 129, 255, 300, 511, 640, 1000, 1337, 1500]))
        kind = str(rng.choice(["thin_plate", "gaussian", "laplace"]))
        P, y = cloud(rng, n)
        p0, p1 = (4.2 + rng.random(), 0.0) if kind == "thin_plate" else (float(rng.uniform(0.5, 2.0)), float(rng.uniform(0.5, 2.0)))
        noise = rng.random() < 0.8
        s2 = rng.uniform(0.02, 0.2, n) if noise else None
        if not noise and kind == "thin_plate" and n > 300:
            s2 = np.full(n, 1e-3)                     # noise-free thin-plate at this size is too ill-conditioned for any tolerance
        normals = bool(rng.random() < 0.5)
        reg = g.GPRegressor(kind, p0, p1, ctx=ctx)
        # the factorisation: the library's own choice (all-FP64 at these sizes) or, for a third of the cases with at least
        # three tile columns, forced INT8-assisted with panels of 1-2 tile columns (ragged sizes, all kernels, per-point
        # noise; an indefinite matrix must fall back to the FP64 factorisation and its trailing-block logic)
        fit_i8 = n > 256 and rng.random() < 0.34
        if fit_i8:
            os.environ.update(GPR_FIT_MODE="int8", GPR_FIT_PANEL=str(int(rng.integers(1, 3))), GPR_FIT_LAST="0")
        try:
            m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2, with_normals=normals)
        finally:
            for k_env in ("GPR_FIT_MODE", "GPR_FIT_PANEL", "GPR_FIT_LAST"):
                os.environ.pop(k_env, None)
        o = oracle.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, kind, p0, p1, factor="llt", with_normals=normals)
        indefinite = o.info != 0
        if indefinite:      # e.g. a nearly noise-free thin-plate matrix with a slightly negative eigenvalue: the reference's
            # pivoted LDLT solves it, and so does the GPU path (trailing pivot block); compare against the LDLT oracle
            o = oracle.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, kind, p0, p1, factor="ldlt", with_normals=normals)
        K = o.get(K=True)["K"]
        w = np.abs(np.linalg.eigvalsh(K))
        cond = float(w.max() / w.min())
        tol_a = max(1e-9, 50 * cond * 2.2e-16)
        tol_f = max(1e-9, 5 * cond * 2.2e-16)
        errs = {"alpha": rel(m.alpha, o.alpha) / tol_a}
        if normals:
            errs["normals"] = float(np.abs(m.get()["normals"] - o.get()["normals"]).max()) / max(1e-8, 100 * cond * 2.2e-16)
        steps = []
        modes = []
        # optional appends
        for _ in range(int(rng.integers(0, 3))):
            k = int(rng.integers(1, 70))
            Pn, yn = cloud(rng, k)
            sn = rng.uniform(0.02, 0.2, k) if s2 is not None else None
            if rng.random() < 0.3:
                reg.reserve(m, m.n + 300)
            try:
                reg.update(m, Pn[:, 0], Pn[:, 1], Pn[:, 2], yn, sn)
            except g.GPRegressionException as e:
                # documented limit: a matrix that is indefinite as a whole (no outlier points to blame) with more than 256
                # points after the first non-positive pivot is reported as NOT_SPD — and the model must be left intact
                if indefinite and e.code == g.GPR_ERR_NOT_SPD and m.n == o.n:
                    steps.append("NOT_SPD(limit)")
                    break
                raise
            o.update(Pn[:, 0], Pn[:, 1], Pn[:, 2], yn, sn)
            steps.append(k)
            errs["alpha_after_append_%d" % len(steps)] = rel(m.alpha, o.alpha) / (2 * tol_a)
        for q in [int(x) for x in rng.choice([1, 2, 5, 8, 9, 63, 130, 1000, 9500, 20000], size=3, replace=False)]:
            Q = rng.uniform(-1.2, 1.2, (q, 3))
            # the form of the variance: the library's own choice, or one forced (INT8 tensor cores / FP64 product / FP64
            # forward substitution; a forced form the model cannot use, e.g. on an indefinite-tail model, falls back inside)
            mode = str(rng.choice(["auto", "auto", "ozaki", "product", "trsm"]))
            if mode == "auto":
                os.environ.pop("GPR_VAR_MODE", None)
            else:
                os.environ["GPR_VAR_MODE"] = mode
            modes.append(mode[0])
            f, v, gr, tx, ty = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, tangent=True)
            os.environ.pop("GPR_VAR_MODE", None)
            f1 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2])
            sub = slice(0, min(q, 400))
            fo, vo, go = o.predict(Q[sub, 0], Q[sub, 1], Q[sub, 2], var=True, grad=True, threads=8)
            k0 = abs(vo).max()
            errs["q%d_mean" % q] = max(rel(f[sub], fo), rel(f1[sub], fo)) / tol_f
            if errs["q%d_mean" % q] > 1.0 and os.environ.get("STRESS_DEBUG"):
                fresh = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2) if not steps or steps == ["NOT_SPD(limit)"] else None
                f2 = reg.evaluate(fresh, Q[:, 0], Q[:, 1], Q[:, 2]) if fresh is not None else f1
                bad_i = np.argsort(-np.abs(f[sub] - fo))[:5]
                print("   DEBUG q=%d: f-vs-oracle %.3e f1-vs-oracle %.3e fresh-model-vs-oracle %.3e | worst idx %s f %s fo %s n_tail %d/%d alpha-dev-vs-host ok" % (
                    q, rel(f[sub], fo), rel(f1[sub], fo), rel(f2[sub], fo), bad_i, f[sub][bad_i], fo[bad_i], m.n_tail, fresh.n_tail if fresh is not None else -1), flush=True)
            errs["q%d_var" % q] = np.abs(v[sub] - vo).max() / (max(1e-7, 50 * cond * 2.2e-16) * max(k0, 1e-12))
            errs["q%d_grad" % q] = rel(gr[sub], go) / tol_f
            N, Tx, Ty = oracle.tangent_basis(gr[sub])
            errs["q%d_tangent" % q] = max(np.abs(tx[sub] - Tx).max(), np.abs(ty[sub] - Ty).max()) / 1e-8
            if not (np.isfinite(f).all() and np.isfinite(v).all() and np.isfinite(gr).all()):
                errs["q%d_finite" % q] = 1e9
        worst = max(errs.values())
        status = "ok" if worst <= 1.0 else "VIOLATION"
        if worst > 1.0:
            bad.append((seed, {k: round(v, 2) for k, v in errs.items() if v > 1.0}))
        print("case %3d seed %4d n=%4d %-10s noise=%d normals=%d fit=%s appends=%s var=%s cond=%.1e tail=%d%s worst=%.3f %s" % (
            case, seed, n, kind, noise, normals, "i8" if fit_i8 else "f64", steps, "".join(modes), cond, m.n_tail, " (indefinite)" if indefinite else "", worst, status), flush=True)
    except Exception as e:     # noqa: BLE001
        bad.append((seed, repr(e)))
        print("case %3d seed %4d EXCEPTION %r" % (case, seed, e), flush=True)
        traceback.print_exc()
print("stress: %d cases, %d violations, %.0f s" % (cases, len(bad), time.time() - t_start))
for b in bad:
    print("  ", b)
sys.exit(1 if bad else 0)

"""INT8-assisted factorisation (launch_cholesky_int8): time and accuracy against the all-FP64 tile Cholesky (GPU box).

For each (panel width, slice count): device time of the factorisation, alpha against the FP64 fit's (both refined), and the
variance of 2 000 queries against the FP64 fit's through the forward-substitution form (which only uses L)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g
W = g.workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P, y, s2 = W.synthetic_cloud(n, seed=0)
Q = W.grid_slab(256, 128, 129)[:2048]
ctx = g.Context()
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
os.environ["GPR_VAR_MODE"] = "trsm"
out = {"n": n, "runs": []}


def fit(mode, panel=None, slices=None, reps=3):
    os.environ["GPR_FIT_MODE"] = mode
    for k, v in (("GPR_FIT_PANEL", panel), ("GPR_FIT_SLICES", slices)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    best = None
    for _ in range(reps):
        m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
        t = ctx.timings()
        if best is None or t["chol_ms"] < best["chol_ms"]:
            best = dict(t)
        alpha = np.array(m.alpha)
        f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        m.close()
    return best, alpha, f, v


t64, a64, f64, v64 = fit("fp64")
out["fp64"] = {"chol_ms": t64["chol_ms"], "fit_total_ms": t64["fit_total_ms"], "solve_ms": t64["solve_ms"]}
print(json.dumps(out["fp64"]), flush=True)
combos = [(16, 7)] if len(sys.argv) > 2 and sys.argv[2] == "quick" else [(4, 7), (6, 7), (8, 7), (10, 7), (12, 7), (16, 7), (24, 7), (8, 6), (8, 8)]
for panel, S in combos:
    t, a, f, v = fit("int8", panel, S)
    r = {"panel_tiles": panel, "slices": S, "chol_ms": t["chol_ms"], "fit_total_ms": t["fit_total_ms"], "fit_int8_slices": t["fit_int8_slices"],
         "alpha_rel_vs_fp64": float(np.abs(a - a64).max() / np.abs(a64).max()),
         "mean_rel_vs_fp64": float(np.abs(f - f64).max() / np.abs(f64).max()),
         "var_rel_vs_fp64": float(np.abs(v - v64).max() / np.abs(v64).max())}
    out["runs"].append(r)
    print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fit_int8_study_n%d.json" % n, "w"), indent=1)

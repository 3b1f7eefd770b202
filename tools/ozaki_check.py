"""GPU check of the INT8 tensor-core variance path (gpr_ozaki.cu), in increasing order of risk; each stage prints a line.
  python tools/ozaki_check.py [stage_max=3]"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g

stage_max = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rng = np.random.default_rng(0)


def expect(A, B, levels, tri):
    S, M, K = A.shape
    A64 = A.astype(np.int64).copy()
    if tri:
        for r in range(M // 128):
            A64[:, r * 128:(r + 1) * 128, 128 * (r + 1):] = 0
    out = np.zeros((levels, M, B.shape[1]), dtype=np.int64)
    for l in range(levels):
        for t in range(S):
            u = l - t
            if 0 <= u < S:
                out[l] += A64[t] @ B[u].astype(np.int64).T
    return out


# stage 1: one slice, one tile, one k-block ... then bigger
for (S, levels, M, N, K, tri) in ((1, 1, 128, 64, 64, False), (1, 1, 128, 80, 256, False), (2, 2, 256, 128, 512, False),
                                  (3, 3, 384, 192, 384, True), (6, 6, 512, 336, 1024, True), (7, 7, 512, 320, 1024, True)):
    A = rng.integers(-64, 65, size=(S, M, K), dtype=np.int8)
    B = rng.integers(-64, 65, size=(S, N, K), dtype=np.int8)
    t0 = time.time()
    if S >= 3:
        A[0, 128:256, 128:] = 0
        A[1, 0:128, 64:128] = 0
    C = g.selftest_i8gemm(A, B, levels, tri, skip_zero_blocks=S >= 3)
    E = expect(A, B, levels, tri)
    bad = int((C.astype(np.int64) != E).sum())
    print("i8gemm S=%d levels=%d M=%d N=%d K=%d tri=%d: mismatches %d of %d (%.2f s)" % (S, levels, M, N, K, tri, bad, E.size, time.time() - t0), flush=True)
    if bad:
        idx = np.argwhere(C.astype(np.int64) != E)[:5]
        for i in idx:
            print("   at", tuple(i), "got", C[tuple(i)], "expected", E[tuple(i)])
        sys.exit(1)
if stage_max < 2:
    sys.exit(0)

# stage 2: the variance through the C-ABI at a moderate size, against the FP64 product form
W = g.workloads
ctx = g.Context()
for n, q in ((2304, 5000), (4096, 20000)):
    P, y, s2 = W.synthetic_cloud(n, seed=21)
    Q = np.random.default_rng(n).uniform(-1.2, 1.2, size=(q, 3))
    reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    os.environ["GPR_VAR_MODE"] = "product"
    f0, v0 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    os.environ["GPR_VAR_MODE"] = "ozaki"
    for base, S in ((254, 5), (254, 6), (254, 7), (128, 6), (128, 7), (128, 8)):
        os.environ["GPR_OZAKI_SLICES"] = str(S)
        os.environ["GPR_OZAKI_BASE"] = str(base)
        f1, v1 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        t = ctx.timings()
        print("n=%d q=%d base=%d slices=%d: var rel diff vs FP64 product %.3e  (var_ms %.2f, int8 kernel %.2f, mean_ms %.2f)"
              % (n, q, base, S, np.abs(v1 - v0).max() / np.abs(v0).max(), t["predict_var_ms"], t["ozaki_ms"], t["predict_mean_ms"]), flush=True)
    del m
if stage_max < 3:
    sys.exit(0)

# stage 3: headline size, timing per batch of 148*128 queries
n = 16384
P, y, s2 = W.synthetic_cloud(n, seed=0)
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
Q = W.grid_slab(256, 128, 129)[:148 * 128]
os.environ["GPR_VAR_MODE"] = "product"
f0, v0 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
f0, v0 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
print("n=16384 product form: var_ms %.2f" % ctx.timings()["predict_var_ms"], flush=True)
os.environ["GPR_VAR_MODE"] = "ozaki"
rows = []
for base, S in ((254, 5), (254, 6), (254, 7), (128, 6), (128, 7), (128, 8)):
    os.environ["GPR_OZAKI_SLICES"] = str(S)
    os.environ["GPR_OZAKI_BASE"] = str(base)
    for rep in range(2):
        f1, v1 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    t = ctx.timings()
    flop = float(n) ** 2 * len(Q)
    rows.append({"n": n, "queries": len(Q), "digit_base": base, "slices": S, "slice_pairs": S * (S + 1) // 2, "tile_n": 80 if S <= 6 else 64,
                 "var_rel_diff_vs_fp64_product": float(np.abs(v1 - v0).max() / np.abs(v0).max()), "var_ms": t["predict_var_ms"],
                 "int8_kernel_ms": t["ozaki_ms"], "fp64_equivalent_tflops": flop / (t["predict_var_ms"] * 1e-3) / 1e12,
                 "int8_tops": 2.0 * (S * (S + 1) // 2) * (n * (n + 128) / 2.0) * len(Q) / (t["ozaki_ms"] * 1e-3) / 1e12 if t["ozaki_ms"] > 0 else None})
    print(rows[-1], flush=True)
import json
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
if os.path.isdir(out):
    json.dump({"what": "INT8 tensor-core variance (gpr_ozaki.cu) at the headline size: accuracy and speed per digit system / slice count; "
                       "FP64 product form (var_tiles_kernel) on the same batch: %.2f ms" % 0.0, "rows": rows}, open(os.path.join(out, "ozaki_table.json"), "w"), indent=1)

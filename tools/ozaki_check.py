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
for (S, levels, M, N, K, tri) in ((1, 1, 128, 64, 64, False), (1, 1, 128, 64, 256, False), (2, 2, 256, 128, 512, False),
                                  (3, 3, 384, 192, 384, True), (7, 7, 512, 320, 1024, True)):
    A = rng.integers(-64, 65, size=(S, M, K), dtype=np.int8)
    B = rng.integers(-64, 65, size=(S, N, K), dtype=np.int8)
    t0 = time.time()
    C = g.selftest_i8gemm(A, B, levels, tri)
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
    for S in (5, 6, 7, 8):
        os.environ["GPR_OZAKI_SLICES"] = str(S)
        f1, v1 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        t = ctx.timings()
        print("n=%d q=%d slices=%d: var rel diff vs FP64 product %.3e  (var_ms %.2f, mean_ms %.2f)"
              % (n, q, S, np.abs(v1 - v0).max() / np.abs(v0).max(), t["predict_var_ms"], t["predict_mean_ms"]), flush=True)
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
for S in (6, 7, 8):
    os.environ["GPR_OZAKI_SLICES"] = str(S)
    for rep in range(2):
        f1, v1 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    t = ctx.timings()
    flop = float(n) ** 2 * len(Q)
    print("n=16384 slices=%d: var rel diff %.3e, var_ms %.2f (%.1f FP64-equivalent TF/s), mean_ms %.2f"
          % (S, np.abs(v1 - v0).max() / np.abs(v0).max(), t["predict_var_ms"], flop / (t["predict_var_ms"] * 1e-3) / 1e12, t["predict_mean_ms"]), flush=True)

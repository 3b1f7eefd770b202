#!/usr/bin/env python
"""CPU study for VERDICT r1 #7: can the variance product V = X K*^T (X = L^-1) run on the INT8 tensor cores (tcgen05
kind::i8, TMEM int32 accumulators) by Ozaki-style slicing, at the variance tolerance of the north star (1e-7 of max v)?

Emulation of csrc/gpr_ozaki.cu in exact integer arithmetic (numpy int64): rows of X are scaled by a power of two, the K*
batch by one power of two, both are cut into S signed digits of base 254 (|digit| <= 127) or base 128 (|digit| <= 64)
(round-to-nearest, remainder carried to the next digit), and V = sum over levels l < S of [sum_{t+u=l} X_t K_u^T] / (F^2 B^l)
with every bracket an exact integer matrix product (what the int8 MMAs accumulate in int32), recombined in float64.
Reports, per (base, S): number of int8 GEMMs, max |dv| / max v against the long-double variance.  The measured table at
the headline size is profiles/ozaki_table_r2.json (tools/ozaki_check.py on the GPU).

  python tools/ozaki_study.py [n=2048] [q=128] > profiles/ozaki_slicing_study_r2.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpr_b200 as g          # host-side workloads only
import oracle


def slices(M, axis, S, base):
    """M = 2^e * sum_t D_t / (F base^t), F = base / 2: the kernel's digit systems (oz_slice in csrc/gpr_ozaki.cu):
    base 128 -> |digit| <= 64, base 254 -> |digit| <= 127."""
    F = base / 2.0
    m = np.abs(M).max(axis=axis, keepdims=True)
    e = np.where(m > 0, np.floor(np.log2(np.maximum(m, 1e-300))) + 1, 0.0)      # 2^e > max, like frexp
    rem = M / np.exp2(e) * F
    out = []
    for t in range(S):
        it = np.rint(rem)
        out.append(it.astype(np.int64))
        rem = (rem - it) * base
    assert max(int(np.abs(o).max()) for o in out) <= F
    return e, out


def slices_fixed(M, e, S, base):
    F = base / 2.0
    rem = M / 2.0 ** e * F
    out = []
    for t in range(S):
        it = np.rint(rem)
        out.append(it.astype(np.int64))
        rem = (rem - it) * base
    return e, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    q = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    W = g.workloads
    P, y, s2 = W.synthetic_cloud(n, seed=0)
    Q = W.grid_slab(64, 30, 31)[:: max(1, 4096 // q)][:q]
    from oracle import oracle as O
    from scipy.linalg import solve_triangular
    m = oracle.blas_fit(P, y, s2, "thin_plate", W.SYNTH_R, 0.0)
    X = solve_triangular(m["L"], np.eye(n), lower=True)
    Ks = O._kern("thin_plate", W.SYNTH_R, 0.0, O._pdist(P, Q))            # n x q
    k0 = W.SYNTH_R ** 3
    ld = np.longdouble
    Vx = (X.astype(ld) @ Ks.astype(ld))
    v_ref = (ld(k0) - (Vx * Vx).sum(axis=0)).astype(np.float64)
    v_f64 = k0 - ((X @ Ks) ** 2).sum(axis=0)
    rows = []
    out = {"n": n, "queries": q, "max_v": float(np.abs(v_ref).max()), "fp64_product_rel_err": float(np.abs(v_f64 - v_ref).max() / np.abs(v_ref).max()),
           "abs_row_sum_X_times_K": float((np.abs(X) @ np.abs(Ks)).max()), "rows": rows}
    for base in (254, 128):
        F = base / 2.0
        for S in (4, 5, 6, 7, 8):
            ex, Xs = slices(X, 1, S, base)
            ek, Ksl = slices(Ks, 0, S, base)
            ek[:] = ek.max()                              # the kernel uses ONE scale for the whole K* batch (2^g >= k(0))
            ek, Ksl = slices_fixed(Ks, float(ek.max()), S, base)
            V = np.zeros((n, q))
            gemms = 0
            for lvl in range(S):                          # levels t + u < S are kept
                acc = np.zeros((n, q), dtype=np.int64)
                for t in range(lvl + 1):
                    acc += Xs[t] @ Ksl[lvl - t]
                    gemms += 1
                assert np.abs(acc).max() < 2 ** 31        # what an int32 TMEM accumulator must hold
                V += acc.astype(np.float64) / (F * F * float(base) ** lvl)
            V *= np.exp2(ex) * np.exp2(ek)
            v = k0 - (V * V).sum(axis=0)
            rows.append({"digit_base": base, "slices": S, "int8_gemms": gemms,
                         "var_rel_err": float(np.abs(v - v_ref).max() / np.abs(v_ref).max())})
            print(rows[-1], file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""CPU study for VERDICT r1 #7: can the variance product V = X K*^T (X = L^-1) run on the INT8 tensor cores (tcgen05
kind::i8, TMEM int32 accumulators) by Ozaki-style slicing, at the variance tolerance of the north star (1e-7 of max v)?

Emulation (exact integer arithmetic, numpy int64): rows of X and columns of K*^T are scaled by a power of two (their
largest magnitude), cut into S signed slices of `bits` bits each (round-to-nearest, remainder carried to the next slice),
and V ~ sum over slice pairs (t, u) with t + u <= level_max of 2^(-bits (t + u + 2)) X_t K_u^T with every slice product an
exact integer matrix product (what an int8 MMA with int32 accumulation computes for k <= 2^31 / 127^2 terms), recombined in
float64.  Reports, per (S, level_max): number of int8 GEMMs, max |dv| / max v against the long-double variance.

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


def slices(M, axis, S, bits):
    """M ~ 2^e * sum_t 2^(-bits (t + 1)) I_t along `axis` scaling; I_t integer with |I_t| <= 2^(bits-1)."""
    e = np.ceil(np.log2(np.abs(M).max(axis=axis, keepdims=True) + 1e-300))
    rem = M / np.exp2(e)                              # in [-1, 1]
    out = []
    for t in range(S):
        scaled = rem * float(1 << bits)
        it = np.rint(scaled)
        out.append(it.astype(np.int64))
        rem = scaled - it                             # in [-0.5, 0.5]
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
    for bits in (7, 6):
        for S in (5, 6, 7, 8, 9):
            ex, Xs = slices(X, 1, S, bits)
            ek, Ksl = slices(Ks, 0, S, bits)
            for level_max in (S - 1, S, 2 * S - 2):
                V = np.zeros((n, q))
                gemms = 0
                for lvl in range(level_max + 1):
                    acc = np.zeros((n, q), dtype=np.int64)
                    for t in range(S):
                        u = lvl - t
                        if 0 <= u < S:
                            acc += Xs[t] @ Ksl[u]
                            gemms += 1
                    V += acc.astype(np.float64) * 2.0 ** (-bits * (lvl + 2))
                V *= np.exp2(ex) * np.exp2(ek)
                v = k0 - (V * V).sum(axis=0)
                rows.append({"bits": bits, "slices": S, "level_max": level_max, "int8_gemms": gemms,
                             "var_rel_err": float(np.abs(v - v_ref).max() / np.abs(v_ref).max())})
                print(rows[-1], file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

// gpr_leaf.cuh — single-CTA leaves of the tile factorisation: Cholesky and triangular inverse of one
// 128x128 tile held column-major in shared memory (pitch PM).  They sit on the critical path of the
// tile-task Cholesky (one diagonal tile per tile column), so they are built for latency:
//   * 16x16 diagonal blocks are factored by ONE warp with a row per lane in registers and the pivot
//     column broadcast by shuffles ("warp-level panel factorisation"), rsqrt instead of sqrt + divide;
//   * the panel below a diagonal block is a right-looking substitution with one row per thread
//     (independent FMAs per column instead of a dependent dot product);
//   * trailing updates inside the tile run on the FP64 tensor pipe (DMMA 8x8x4 on 8x8 sub-tiles), and
//     the update that the next diagonal block does not need is overlapped with that block's
//     factorisation (look-ahead: warp 0 factors while warps 1..7 update);
//   * the triangular inverse is computed IN PLACE bottom-up: 16x16 diagonal inverses by the 8 warps in
//     parallel, then X21 = -X22 (L21 X11) for block sizes 16, 32, 64 as DMMA block products.
#pragma once
#include "gpr_mma.cuh"

namespace gpr {

// dst(s x s) = sign * A(s x s) * B(s x s); all column-major in shared memory (element (r,c) at
// base[c*pitch + r]).  A_LOWER: A is lower triangular (k <= m); B_LOWER: B is lower triangular (k >= n);
// the strict upper part of diagonal 8x8 sub-tiles must hold zeros.  The (s/8)^2 output tiles are shared
// by the nw warps whose local index is wloc.
template <bool A_LOWER, bool B_LOWER>
__device__ __forceinline__ void block_mm(double* dst, int pd, const double* A, int pa, const double* B, int pb,
                                         int s, double sign, int wloc, int nw, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int nt = s >> 3;
    for (int idx = wloc; idx < nt * nt; idx += nw) {
        const int tm = idx % nt, tn = idx / nt;
        const int m0 = 8 * tm, n0 = 8 * tn;
        const int kb = B_LOWER ? n0 : 0;
        const int ke = A_LOWER ? m0 + 8 : s;
        double c0 = 0.0, c1 = 0.0;
        for (int k0 = kb; k0 < ke; k0 += 4) dmma(c0, c1, A[(k0 + t) * pa + m0 + g], B[(n0 + g) * pb + k0 + t]);
        dst[(n0 + 2 * t) * pd + m0 + g] = sign * c0;
        dst[(n0 + 2 * t + 1) * pd + m0 + g] = sign * c1;
    }
}

// Trailing update of one 8x8 tile: T[r0+..][q0+..] -= sum_k P[r][k] P[q][k], the panel P being the 16
// columns [pc, pc+16).  One warp, 4 DMMA.
__device__ __forceinline__ void leaf_update_tile(double* S, int pc, int r0, int q0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int k = pc + 4 * s + t;
        dmma(c0, c1, S[k * PM + r0 + g], S[k * PM + q0 + g]);
    }
    S[(q0 + 2 * t) * PM + r0 + g] -= c0;
    S[(q0 + 2 * t + 1) * PM + r0 + g] -= c1;
}

// All lower tiles (ti >= tj, counted from row/column `base`) of the trailing update of panel [pc, pc+16)
// EXCEPT the three tiles of the leading 16x16 block (ti < 2), shared by nw warps.
__device__ __forceinline__ void leaf_update_rest(double* S, int pc, int base, int wloc, int nw, int lane) {
    const int nt8 = (TB - base) >> 3;
    int ti = 2, tj = 0, idx = 0;
    for (int want = wloc;; want += nw) {
        while (idx < want) { ++idx; if (++tj > ti) { ++ti; tj = 0; } }
        if (ti >= nt8) break;
        leaf_update_tile(S, pc, base + 8 * ti, base + 8 * tj, lane);
    }
}

// Cholesky of the 16x16 diagonal block at (c0,c0) by one warp.  Lane l (< 16) owns row l.
// sinv[c0 + c] receives 1 / L[c][c].
// The column loops are unrolled by template recursion: with plain nested "#pragma unroll" loops whose
// inner bound depends on the outer index the compiler kept a[] in local memory (LDL/STL per step).
template <int C, int CC>
struct Potrf16Update {
    static __device__ __forceinline__ void run(double (&a)[16], int l) {
        const double lcc = __shfl_sync(0xffffffffu, a[C], CC);
        const double upd = fma(-a[C], lcc, a[CC]);
        a[CC] = (l >= CC) ? upd : a[CC];
        Potrf16Update<C, CC + 1>::run(a, l);
    }
};
template <int C>
struct Potrf16Update<C, 16> {
    static __device__ __forceinline__ void run(double (&)[16], int) {}
};

// pivot -> (1/sqrt(p), sqrt(p)): rsqrt (1 ulp) and one Newton step for the square root.
__device__ __forceinline__ void potrf16_pivot(double p, int col, int& bad, double& inv, double& d) {
    if (!(p > 0.0)) { bad = min(bad, col); p = 1.0; }
    inv = rsqrt(p);
    d = p * inv;
    d = fma(0.5 * inv, fma(-d, d, p), d);
}

// Column C with its pivot already reduced to (inv, d).  The next pivot only needs the update of entry
// C+1, so that update, the broadcast of the new pivot and its rsqrt are issued first and overlap with
// the remaining updates of column C (software pipelining of the only long-latency chain).
template <int C>
struct Potrf16Col {
    static __device__ __forceinline__ void run(double (&a)[16], int l, int lane, double* sinv, int c0, int& bad,
                                               double inv, double d) {
        a[C] = (l == C) ? d : ((l > C) ? a[C] * inv : a[C]);
        if (lane == C) sinv[c0 + C] = inv;
        const double l1 = __shfl_sync(0xffffffffu, a[C], C + 1);
        const double u1 = fma(-a[C], l1, a[C + 1]);
        a[C + 1] = (l >= C + 1) ? u1 : a[C + 1];
        double inv1, d1;
        potrf16_pivot(__shfl_sync(0xffffffffu, a[C + 1], C + 1), c0 + C + 1, bad, inv1, d1);
        Potrf16Update<C, C + 2>::run(a, l);
        Potrf16Col<C + 1>::run(a, l, lane, sinv, c0, bad, inv1, d1);
    }
};
template <>
struct Potrf16Col<15> {
    static __device__ __forceinline__ void run(double (&a)[16], int l, int lane, double* sinv, int c0, int&,
                                               double inv, double d) {
        a[15] = (l == 15) ? d : a[15];
        if (lane == 15) sinv[c0 + 15] = inv;
    }
};
__device__ __forceinline__ void leaf_potrf16(double* S, int c0, double* sinv, int* s_fail, int lane) {
    const int l = lane & 15;
    double a[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) a[c] = S[(c0 + c) * PM + c0 + l];
    int bad = 1 << 20;
    double inv0, d0;
    potrf16_pivot(__shfl_sync(0xffffffffu, a[0], 0), c0, bad, inv0, d0);
    Potrf16Col<0>::run(a, l, lane, sinv, c0, bad, inv0, d0);
    if (bad < TB && lane == 0) atomicMin(s_fail, bad);
    if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; ++c) S[(c0 + c) * PM + c0 + l] = (l >= c) ? a[c] : 0.0;
    }
}

// Column c (one per lane) of the inverse of a 16x16 lower-triangular block, rows unrolled by recursion.
template <int R, int K>
struct Inv16Dot {
    static __device__ __forceinline__ void run(const double* Lrow, const double (&x)[16], double& s0, double& s1) {
        const double lrk = Lrow[K * PM];                       // L[r][k], same address for all lanes
        if (K & 1) s1 = fma(lrk, x[K], s1); else s0 = fma(lrk, x[K], s0);
        Inv16Dot<R, K + 1>::run(Lrow, x, s0, s1);
    }
};
template <int R>
struct Inv16Dot<R, R> {
    static __device__ __forceinline__ void run(const double*, const double (&)[16], double&, double&) {}
};
template <int R>
struct Inv16Row {
    static __device__ __forceinline__ void run(const double* Lblk, const double* sinv, double (&x)[16], int c) {
        double s0 = 0.0, s1 = 0.0;
        Inv16Dot<R, 0>::run(Lblk + R, x, s0, s1);
        const double inv = sinv[R];
        x[R] = (R == c) ? inv : ((R > c) ? -inv * (s0 + s1) : 0.0);
        Inv16Row<R + 1>::run(Lblk, sinv, x, c);
    }
};
template <>
struct Inv16Row<16> {
    static __device__ __forceinline__ void run(const double*, const double*, double (&)[16], int) {}
};

// ---------------------------------------------------------------------------------------------
// Cholesky of a 128x128 tile in shared memory.  On exit the lower triangle holds L, the strict upper
// part of the diagonal 16x16 blocks is zero (the rest of the upper triangle is NOT cleaned: callers
// write zeros for r < c themselves), sinv[0..127] = 1 / diag(L).
// *s_fail receives the smallest failing column (pivot <= 0 or NaN), or stays >= 128.
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ void potrf128_smem(double* S, double* sinv, int* s_fail, long long* prof = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tA = 0, tAw = 0, tB = 0, tC = 0, tC2 = 0, t0 = 0;
    for (int b = 0; b < 8; ++b) {
        const int c0 = 16 * b;
        // (A) diagonal block on warp 0, overlapped with the part of the previous panel's trailing update
        //     that this block column does not depend on (columns >= c0 + 16), on warps 1..7.
        if (prof) t0 = clock64();
        if (warp == 0) leaf_potrf16(S, c0, sinv, s_fail, lane);
        else if (b > 0) leaf_update_rest(S, c0 - 16, c0, warp - 1, 7, lane);
        if (prof) { long long t = clock64(); if (warp == 0) tA += t - t0; else tC2 += t - t0; t0 = t; }
        __syncthreads();
        if (prof) { long long t = clock64(); tAw += t - t0; t0 = t; }
        if (b == 7) break;
        // (B) panel below the diagonal block: one thread per row, right-looking substitution
        if (tid < TB && tid >= c0 + 16) {
            const int r = tid;
            double x[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = S[(c0 + c) * PM + r];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                x[c] *= sinv[c0 + c];
#pragma unroll
                for (int cc = c + 1; cc < 16; ++cc) x[cc] = fma(-x[c], S[(c0 + c) * PM + c0 + cc], x[cc]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) S[(c0 + c) * PM + r] = x[c];
        }
        __syncthreads();
        if (prof) { long long t = clock64(); tB += t - t0; t0 = t; }
        // (C1) update of the next diagonal 16x16 block only (3 tiles); everything else of this panel's
        //      trailing update runs on warps 1..7 while warp 0 factors that block (phase A above).
        if (warp < 3) leaf_update_tile(S, c0, c0 + 16 + 8 * ((warp + 1) >> 1), c0 + 16 + 8 * (warp >> 1), lane);
        __syncthreads();
        if (prof) { long long t = clock64(); tC += t - t0; t0 = t; }
    }
    if (prof) {
        if (tid == 0) { prof[0] = tA; prof[1] = tAw; prof[2] = tB; prof[3] = tC; }
        if (tid == 32) prof[4] = tC2;
        if (tid == 127) prof[5] = tB;      // a thread that does panel work in every block
    }
}

// ---------------------------------------------------------------------------------------------
// X = L^-1 in place for the 128x128 lower-triangular tile in shared memory (as left by potrf128_smem).
// tmp: scratch of at least 64*68 doubles.  sinv: 1/diag(L).  On exit the lower triangle of S holds X
// (upper triangle: zero inside diagonal 16x16 blocks, garbage elsewhere).
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ void trinv128_smem(double* S, const double* sinv, double* tmp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        // 16x16 diagonal inverses, one block per warp; lane c (< 16) owns column c of X.
        const int c0 = 16 * warp, c = lane & 15;
        double x[16];
        Inv16Row<0>::run(S + c0 * PM + c0, sinv + c0, x, c);
        __syncwarp();
        if (lane < 16) {
#pragma unroll
            for (int r = 0; r < 16; ++r) S[(c0 + c) * PM + c0 + r] = x[r];
        }
    }
    __syncthreads();
    for (int s = 16; s <= 64; s <<= 1) {
        const int npairs = TB / (2 * s);
        const int nwp = 8 / npairs;
        const int pair = warp / nwp, wloc = warp % nwp;
        const int o = pair * 2 * s;
        const int pw = s + 4;
        double* X11 = S + o * PM + o;
        double* X22 = S + (o + s) * PM + (o + s);
        double* L21 = S + o * PM + (o + s);
        double* W = tmp + pair * s * pw;
        block_mm<false, true>(W, pw, L21, PM, X11, PM, s, 1.0, wloc, nwp, lane);       // W = L21 X11
        __syncthreads();
        block_mm<true, false>(L21, PM, X22, PM, W, pw, s, -1.0, wloc, nwp, lane);      // X21 = -X22 W
        __syncthreads();
    }
}

// Copy the lower triangle of a shared-memory tile (pitch PM) to a column-major global tile, writing
// zeros above the diagonal.
__device__ __forceinline__ void store_lower_tile(const double* S, double* G, size_t ld) {
    for (int idx = threadIdx.x; idx < TB * TB; idx += NTHREADS) {
        const int r = idx & (TB - 1), c = idx >> 7;
        G[(size_t)c * ld + r] = (r >= c) ? S[c * PM + r] : 0.0;
    }
}

}  // namespace gpr

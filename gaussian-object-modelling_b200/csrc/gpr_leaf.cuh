// gpr_leaf.cuh — single-CTA leaves of the tile factorisation: Cholesky and triangular inverse of one
// 128x128 tile held in shared memory ("warp-level panel factorisation": the 16x16 diagonal blocks are
// factored by one warp with register rows and shuffles; trailing updates inside the tile use DMMA).
#pragma once
#include "gpr_mma.cuh"

namespace gpr {

// ---------------------------------------------------------------------------------------------
// Leaf: Cholesky of a 128x128 tile held column-major in shared memory (pitch PM).
// On exit the lower triangle holds L and the strict upper triangle is zero.
// *s_fail receives the smallest failing column (pivot <= 0 or NaN), or stays >= 128.
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ void potrf128_smem(double* S, int* s_fail) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int b = 0; b < 8; ++b) {
        const int c0 = 16 * b;
        // (A) 16x16 diagonal block: one warp, lane l holds row l in registers, columns by shuffles.
        if (warp == 0) {
            const int l = lane & 15;
            double a[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) a[c] = S[(c0 + c) * PM + c0 + l];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                double p = __shfl_sync(0xffffffffu, a[c], c);
                if (!(p > 0.0)) {
                    if (lane == 0) atomicMin(s_fail, c0 + c);
                    p = 1.0;
                }
                double d = sqrt(p);
                double inv = 1.0 / d;
                if (l == c) a[c] = d;
                else if (l > c) a[c] *= inv;
#pragma unroll
                for (int cc = c + 1; cc < 16; ++cc) {
                    double lcc = __shfl_sync(0xffffffffu, a[c], cc);
                    if (l >= cc) a[cc] -= a[c] * lcc;
                }
            }
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 16; ++c) S[(c0 + c) * PM + c0 + l] = (l >= c) ? a[c] : 0.0;
            }
        }
        __syncthreads();
        if (b == 7) break;
        // (B) panel below the diagonal block: one thread per row, forward substitution with L_dd.
        if (tid < TB && tid >= c0 + 16) {
            const int r = tid;
            double x[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = S[(c0 + c) * PM + r];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                double s = x[c];
#pragma unroll
                for (int k = 0; k < c; ++k) s -= x[k] * S[(c0 + k) * PM + c0 + c];
                x[c] = s / S[(c0 + c) * PM + c0 + c];
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) S[(c0 + c) * PM + r] = x[c];
        }
        __syncthreads();
        // (C) trailing update T[r][cc] -= sum_k P[r][k] P[cc][k] on 8x8 tiles (lower tiles only), DMMA.
        const int base = c0 + 16;
        const int nt8 = (TB - base) >> 3;
        for (int idx = warp; idx < nt8 * nt8; idx += 8) {
            const int ti = idx / nt8, tj = idx - ti * nt8;
            if (ti < tj) continue;
            const int r0 = base + 8 * ti, q0 = base + 8 * tj;
            double c0v = 0.0, c1v = 0.0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int k = c0 + 4 * s + t;
                dmma(c0v, c1v, S[k * PM + r0 + g], S[k * PM + q0 + g]);
            }
            S[(q0 + 2 * t) * PM + r0 + g] -= c0v;
            S[(q0 + 2 * t + 1) * PM + r0 + g] -= c1v;
        }
        __syncthreads();
    }
    // zero the strict upper triangle outside the diagonal 16x16 blocks
    for (int idx = tid; idx < TB * TB; idx += NTHREADS) {
        const int r = idx & (TB - 1), c = idx >> 7;
        if ((r >> 4) < (c >> 4)) S[c * PM + r] = 0.0;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Leaf: X = L^-1 for the 128x128 lower-triangular tile in shared memory; X is written to global
// memory `out` (column-major, ld 128, strict upper triangle zero).  The strict upper triangle of S
// is used as scratch (X^T) and is left dirty; the lower triangle (L) is preserved.
// Two adjacent lanes share one column c and split the k-sum (even / odd k).
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ void trinv128_smem(double* S, double* __restrict__ out) {
    const int tid = threadIdx.x;
    const int c = tid >> 1, h = tid & 1;
    const int cw = (tid & ~31) >> 1;   // smallest column handled by this warp
    double xdiag = 0.0;
    for (int r = 0; r < TB; ++r) {
        // x_r(c) = (delta_rc - sum_{k=c}^{r-1} L[r][k] x_k(c)) / L[r][r]   for r >= c
        double s = 0.0;
        if (r >= cw) {
            for (int k = cw + h; k < r; k += 2) {
                if (k >= c) {
                    double xk = (k == c) ? xdiag : S[k * PM + c];   // X[k][c] kept at upper position (c,k)
                    s = fma(-S[k * PM + r], xk, s);
                }
            }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (r >= c) {
            double x = ((r == c ? 1.0 : 0.0) + s) / S[r * PM + r];
            if (r == c) xdiag = x;
            else if (h == 0) S[r * PM + c] = x;
        }
        __syncwarp();
    }
    __syncthreads();
    for (int idx = tid; idx < TB * TB; idx += NTHREADS) {
        const int r = idx & (TB - 1), cc = idx >> 7;
        double v = 0.0;
        if (r > cc) v = S[r * PM + cc];
        else if (r == cc) v = 1.0 / S[r * PM + r];
        out[(size_t)cc * TB + r] = v;
    }
    __syncthreads();
}

}  // namespace gpr

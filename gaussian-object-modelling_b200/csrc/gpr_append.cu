// gpr_append.cu — K5: incremental append of k <= 32 training points to a fitted model (rank-k row append of
// the Cholesky factor AND of its inverse), instead of the reference's full refit.
//
// Replaces GPRegressor::update<>() (/root/reference/include/gp_regression/gp_regressor.hpp:367-479), which
// grows Kpp by conservativeResize (:442-452) and then calls LDLT::compute on the whole matrix again
// (:457-459).  The only incremental precedent in the reference is the row append of its unused second
// library (include/gp/GaussianProcess.h:340-374).
//
// With K' = [[K, P], [P^T, C]], L = chol(K), X = L^-1 (kept resident for the variance path):
//     B   = X P                      (n x k)   new rows of L are B^T
//     S   = C - B^T B                (k x k)   Schur complement
//     L22 = chol(S),  W = L22^-1
//     L'  = [[L, 0], [B^T, L22]]
//     X'  = [[X, 0], [-W (B^T X), W]]
// The two n^2*k products read the triangle of X once each (4 n^2 bytes): they are HBM/L2-bandwidth
// bound skinny products.  The 128x128 DMMA tile engine would waste 3/4 of its flops on a 32-wide operand and
// serialise on 64 long tasks, so they have their own kernel: m8n8k4 DMMA fragments loaded straight from global
// memory, split over k (skinny_dmma_kernel); an FMA variant on 32x32 register tiles serves the tail build.
// Every reduction has a fixed order: the append is bit-reproducible.
#include "gpr_mma.cuh"
#include "gpr_leaf.cuh"
#include "gpr_kernels.h"

namespace gpr {

constexpr int AK = 32;      // slab width: new points handled per pass
constexpr int KC = 64;      // k extent of one shared-memory chunk of the skinny products
constexpr int APITCH = 65;  // pitch of the k-contiguous A chunk (odd: conflict-free column reads)

// ---------------------------------------------------------------------------------------------
// (1) cross-covariance of the new points against the old ones, and among themselves.
//     Pn[c*32 + a] = k(|p_c - new_a|)  (c < n0, a < k; 0 for a >= k)
//     S0[a*32 + b] = k(|new_a - new_b|) + [a==b] sigma2_a   (identity for a or b >= k)
// Same arithmetic as cov_build_kernel (dist_exact / kern_value_exact): an appended model has bit-identical
// covariance entries to a refitted one.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) append_panel_kernel(const double* x, const double* y, const double* z,
                                                           const double* sigma2, int n0, int t0, int k, double* Pn,
                                                           double* S0, KernParams kp) {
    __shared__ double nx[AK], ny[AK], nz[AK];
    if (threadIdx.x < AK) {
        const bool real = threadIdx.x < k;
        nx[threadIdx.x] = real ? x[t0 + threadIdx.x] : 0.0;     // the new points sit at [t0, t0 + k)
        ny[threadIdx.x] = real ? y[t0 + threadIdx.x] : 0.0;
        nz[threadIdx.x] = real ? z[t0 + threadIdx.x] : 0.0;
    }
    __syncthreads();
    const int idx = blockIdx.x * 256 + threadIdx.x;          // (c, a) with a fastest
    const int c = idx >> 5, a = idx & 31;
    if (c < n0) {
        double v = 0.0;
        if (a < k) v = kern_value_exact(kp, dist_exact(x[c], y[c], z[c], nx[a], ny[a], nz[a]));
        Pn[idx] = v;
    }
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < AK * AK; e += 256) {
            const int a2 = e >> 5, b2 = e & 31;
            double v = (a2 == b2) ? 1.0 : 0.0;
            if (a2 < k && b2 < k) {
                v = kern_value_exact(kp, dist_exact(nx[a2], ny[a2], nz[a2], nx[b2], ny[b2], nz[b2]));
                if (a2 == b2) v = __dadd_rn(v, sigma2[t0 + a2]);       // gp_regressor.hpp:449-452
            }
            S0[e] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// (2)/(5) skinny triangular products against X = L^-1 (column-major, leading dimension ld, lower part).
//   MODE 0:  OUT[r][a] = sum_{c <= r} X[r][c] * Bm[c][a]          (B = X * Pn)        block rows r0 = 32*blockIdx.x
//   MODE 1:  OUT[c][b] = sum_{r >= c, r < n0} X[r][c] * Bm[r][b]   (G = X^T * B)       block columns c0 = 32*blockIdx.x
//   MODE 2:  OUT[m][a] = sum_{k < kdim} A[k*ld + m] * Bm[k][a]     (plain skinny product, no triangle; rows m < n0)
// OUT and Bm are n0 x 32 row-major.  One CTA = one 32x32 output block; 256 threads = 4 k-phases x (8 x 8)
// threads with a 4x4 register tile each (rows tm + 8i, columns 4tn + j); the k-phases are summed in a
// fixed order at the end.  Chunks of 64 k are double-buffered through registers.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) skinny_tri_kernel(const double* __restrict__ X, size_t ld, int n0, int kdim,
                                                         const double* __restrict__ Bm, double* __restrict__ OUT) {
    __shared__ __align__(16) double As[4 * AK * AK];      // MODE 0: [k][m] pitch 32;  MODE 1: [m][k] pitch 65; then the k-phase sums
    __shared__ __align__(16) double Bs[KC * AK];          // [k][a]
    const int tid = threadIdx.x;
    // heavy blocks first: MODE 0 rows near n0 have the longest k range, MODE 1 columns near 0
    const int nblk = (n0 + AK - 1) / AK;
    const int blk = MODE == 0 ? (nblk - 1 - blockIdx.x) : blockIdx.x;
    const int m0 = blk * AK;
    const int kbeg = MODE == 1 ? m0 : 0;
    const int kend = MODE == 0 ? min(m0 + AK, n0) : (MODE == 1 ? n0 : kdim);
    const int kg = tid >> 6, t64 = tid & 63, tm = t64 & 7, tn = t64 >> 3;

    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    double ra[8], rb[8];
    auto load_chunk = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int idx = tid + 256 * j;
            int mm, kk;
            if (MODE != 1) { kk = idx >> 5; mm = idx & 31; } else { mm = idx >> 6; kk = idx & 63; }
            const int kglob = k0 + kk, mglob = m0 + mm;
            double v = 0.0;
            if (kglob < kend && mglob < n0) {
                if (MODE == 0) { if (kglob <= mglob) v = __ldg(X + (size_t)kglob * ld + mglob); }
                else if (MODE == 1) { if (kglob >= mglob) v = __ldg(X + (size_t)mglob * ld + kglob); }
                else v = __ldg(X + (size_t)kglob * ld + mglob);
            }
            ra[j] = v;
            const int kb = k0 + (idx >> 5);
            rb[j] = kb < kend ? __ldg(Bm + (size_t)kb * AK + (idx & 31)) : 0.0;
        }
    };
    auto store_chunk = [&]() {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int idx = tid + 256 * j;
            if (MODE != 1) As[idx] = ra[j];                       // [k][m], pitch 32
            else As[(idx >> 6) * APITCH + (idx & 63)] = ra[j];    // [m][k], pitch 65
            Bs[idx] = rb[j];
        }
    };

    load_chunk(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += KC) {
        __syncthreads();
        store_chunk();
        __syncthreads();
        if (k0 + KC < kend) load_chunk(k0 + KC);
#pragma unroll 4
        for (int kk = kg; kk < KC; kk += 4) {
            double a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = MODE != 1 ? As[kk * AK + tm + 8 * i] : As[(tm + 8 * i) * APITCH + kk];
            const double2 b01 = *reinterpret_cast<const double2*>(Bs + kk * AK + 4 * tn);
            const double2 b23 = *reinterpret_cast<const double2*>(Bs + kk * AK + 4 * tn + 2);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fma(a[i], b01.x, acc[i][0]); acc[i][1] = fma(a[i], b01.y, acc[i][1]);
                acc[i][2] = fma(a[i], b23.x, acc[i][2]); acc[i][3] = fma(a[i], b23.y, acc[i][3]);
            }
        }
    }
    // sum the four k-phases in a fixed order through shared memory (As is reused)
    __syncthreads();
    double* red = As;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[kg * 1024 + (tm + 8 * i) * AK + 4 * tn + j] = acc[i][j];
    __syncthreads();
    for (int e = tid; e < AK * AK; e += 256) {
        const int mm = e >> 5;
        if (m0 + mm < n0) OUT[(size_t)(m0 + mm) * AK + (e & 31)] = ((red[e] + red[1024 + e]) + red[2048 + e]) + red[3072 + e];
    }
}

// ---------------------------------------------------------------------------------------------
// (2') The same two products on the FP64 tensor pipe, split over k (used by the append slab; the FMA kernel
// above stays for the few calls of the indefinite-tail build).  A 32-wide operand fits the m8n8k4 DMMA shape
// exactly (4 n-tiles), so no flops are wasted.
//   CTA  = 128 rows x one k-span of SPAN (512) columns, 8 warps, warp = 16 rows (two 8-row groups sharing the panel
//          fragments) over the whole span;
//   X    fragments straight from global memory into registers (8 rows x 4 k = sector-exact pieces), each register
//          reloaded for the next 32-k chunk right after its last use: 16 loads per lane always in flight;
//   panel chunk [32 k][32] through shared memory (cp.async, two stages, pitch 40: conflict-free fragment reads) -
//          fragment loads from global memory would touch 4 lines each and saturate the L1 pipe (measured: 94 %);
//   part[span][m][a] partial sums, then skinny_finish_kernel adds the spans of each row in ascending order.
// ---------------------------------------------------------------------------------------------
constexpr int SPAN = 512;       // small uniform work items: ~650 CTAs at n = 8 250 for 296 resident slots
constexpr int SROWS = 128;      // rows per CTA
constexpr int SKC = 32;         // k per shared-memory chunk of the panel
constexpr int SPB = 40;         // panel pitch in shared memory
template <int MODE>
__global__ void __launch_bounds__(256, 2) skinny_dmma_kernel(const double* __restrict__ X, size_t ld, int n0,
                                                             const double* __restrict__ Bm, double* __restrict__ part) {
    __shared__ __align__(16) double Bs[2][SKC * SPB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int nblk = (n0 + SROWS - 1) / SROWS;
    const int blk = MODE == 0 ? (nblk - 1 - blockIdx.x) : blockIdx.x;      // heavy blocks first
    const int span = blockIdx.y;
    const int rbeg = blk * SROWS, rend = min(rbeg + SROWS, n0);
    const int kbeg = max(MODE == 0 ? 0 : rbeg, span * SPAN);                // multiples of 32
    const int kend = min(MODE == 0 ? rend : n0, (span + 1) * SPAN);
    if (kbeg >= kend) return;
    const int m0 = rbeg + 16 * warp + g;                                     // row of group 0; group 1 is m0 + 8

    auto load_a = [&](int k, int m) -> double {
        if (k >= kend || m >= n0) return 0.0;
        if (MODE == 0) return k <= m ? __ldcs(X + (size_t)k * ld + m) : 0.0;
        return k >= m ? __ldcs(X + (size_t)m * ld + k) : 0.0;
    };
    auto stage_b = [&](int kc, int st) {                                     // 32 x 32 doubles = 512 x 16 bytes
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid + 256 * e, kk = idx >> 4, c2 = (idx & 15) * 2;
            double* dst = &Bs[st][kk * SPB + c2];
            if (kc + kk < kend) cp_async16(dst, Bm + (size_t)(kc + kk) * AK + c2);
            else { dst[0] = 0.0; dst[1] = 0.0; }
        }
        cp_async_commit();
    };

    double acc[2][4][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[r][j][0] = 0.0; acc[r][j][1] = 0.0; }
    double a[8][2];
    stage_b(kbeg, 0);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u][0] = load_a(kbeg + 4 * u + t, m0); a[u][1] = load_a(kbeg + 4 * u + t, m0 + 8); }
    int st = 0;
    for (int kc = kbeg; kc < kend; kc += SKC, st ^= 1) {
        const bool more = kc + SKC < kend;
        if (more) stage_b(kc + SKC, st ^ 1);
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        const double* bs = &Bs[st][t * SPB + g];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bs[4 * u * SPB + 8 * j];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                dmma(acc[0][j][0], acc[0][j][1], a[u][0], b[j]);
                dmma(acc[1][j][0], acc[1][j][1], a[u][1], b[j]);
            }
            if (more) { a[u][0] = load_a(kc + SKC + 4 * u + t, m0); a[u][1] = load_a(kc + SKC + 4 * u + t, m0 + 8); }
        }
        __syncthreads();
    }
    // C fragment: row g, columns 8j + 2t, 8j + 2t + 1
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int m = m0 + 8 * r;
        if (m < n0) {
            double* out = part + ((size_t)span * n0 + m) * AK;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(out + 8 * j + 2 * t) = make_double2(acc[r][j][0], acc[r][j][1]);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) skinny_finish_kernel(const double* __restrict__ part, int n0, double* __restrict__ OUT) {
    const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= (size_t)n0 * AK) return;
    const int m = (int)(e >> 5);
    const int rbeg = (m / SROWS) * SROWS, rend = min(rbeg + SROWS, n0);
    const int s0 = MODE == 0 ? 0 : rbeg / SPAN;
    const int s1 = MODE == 0 ? (rend - 1) / SPAN : (n0 - 1) / SPAN;
    double v = 0.0;
    for (int sp = s0; sp <= s1; ++sp) v += part[(size_t)sp * n0 * AK + e];
    OUT[e] = v;
}

// ---------------------------------------------------------------------------------------------
// (3) partial Gram matrices of B: part[blk][a*32 + b] = sum_{r in block} B[r][a] B[r][b], 256 rows per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int GRAM_ROWS = 256;
__global__ void __launch_bounds__(256) append_gram_kernel(const double* __restrict__ B, int n0, double* __restrict__ part) {
    __shared__ double Bs[64 * AK];
    const int tid = threadIdx.x, a = tid & 31, b0 = (tid >> 5) * 4;
    const int r0 = blockIdx.x * GRAM_ROWS;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int rc = 0; rc < GRAM_ROWS; rc += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * AK; e += 256) {
            const int r = r0 + rc + (e >> 5);
            Bs[e] = r < n0 ? B[(size_t)r * AK + (e & 31)] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 64; ++r) {
            const double va = Bs[r * AK + a];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fma(va, Bs[r * AK + b0 + j], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) part[(size_t)blockIdx.x * AK * AK + a * AK + b0 + j] = acc[j];
}

// ---------------------------------------------------------------------------------------------
// (4) S = S0 - sum_blk part[blk] (fixed order, 256 threads);  L22 = chol(S), W = L22^-1 by ONE warp with a row per
// lane in registers (right-looking Cholesky with shuffle broadcasts, then row r of W from w_r L22 = e_r; the pivot
// arithmetic is the tile leaf's potrf16_pivot).  out22: [0,1024) L22, [1024,2048) W, both row-major [a*32 + b].
// *flag = 0, or 1 + global index of the first non-positive pivot (sticky over the slabs of one append: once set,
// the later kernels of the append do nothing).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) append_leaf_kernel(const double* __restrict__ S0, const double* __restrict__ part,
                                                          int nparts, double* __restrict__ out22, int* flag, int n0) {
    __shared__ double S[AK * (AK + 1)];
    const int tid = threadIdx.x;
    if (*flag != 0) return;                  // an earlier slab of this append failed: the flag is sticky
    for (int e = tid; e < AK * AK; e += 256) {
        double s = 0.0;
        for (int p = 0; p < nparts; ++p) s += part[(size_t)p * AK * AK + e];
        S[(e >> 5) * (AK + 1) + (e & 31)] = S0[e] - s;
    }
    __syncthreads();
    if (tid >= 32) return;
    const int lane = tid;
    const unsigned full = 0xffffffffu;
    double a[AK], dinv[AK];
#pragma unroll
    for (int c = 0; c < AK; ++c) a[c] = S[lane * (AK + 1) + c];
    int bad = 1 << 20;
#pragma unroll
    for (int j = 0; j < AK; ++j) {
        double inv, d;
        potrf16_pivot(__shfl_sync(full, a[j], j), j, bad, inv, d);
        dinv[j] = inv;
        a[j] = (lane == j) ? d : ((lane > j) ? a[j] * inv : a[j]);
#pragma unroll
        for (int c = j + 1; c < AK; ++c) {
            const double lc = __shfl_sync(full, a[j], c);
            if (lane >= c) a[c] = fma(-a[j], lc, a[c]);
        }
    }
    if (bad < AK) {
        if (lane == 0) *flag = n0 + bad + 1;
        return;
    }
#pragma unroll
    for (int c = 0; c < AK; ++c) out22[lane * AK + c] = (c <= lane) ? a[c] : 0.0;
    double w[AK];
#pragma unroll
    for (int c = AK - 1; c >= 0; --c) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = c + 1; k < AK; ++k) {
            const double lkc = __shfl_sync(full, a[c], k);            // L22[k][c]; w[k] is zero for k > lane
            if (k & 1) s1 = fma(w[k], lkc, s1); else s0 = fma(w[k], lkc, s0);
        }
        w[c] = (c == lane) ? dinv[c] : ((c < lane) ? -(s0 + s1) * dinv[c] : 0.0);
    }
#pragma unroll
    for (int c = 0; c < AK; ++c) out22[AK * AK + lane * AK + c] = w[c];
}

// ---------------------------------------------------------------------------------------------
// (6) scatter the new rows into L and X (nothing is written when the slab failed):
//   L[n0+a][c] = B[c][a],  X[n0+a][c] = -sum_{b<=a} W[a][b] G[c][b]   (c < n0)
//   L[n0+a][n0+b] = L22[a][b],  X[n0+a][n0+b] = W[a][b]                (b <= a; zero above the diagonal)
// then refresh the inverse diagonal tiles: Dinv_t = X_tt for the tiles the new rows touch.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) append_scatter_kernel(const double* __restrict__ B, const double* __restrict__ G,
                                                             const double* __restrict__ out22, const int* flag, int n0,
                                                             int k, double* L, double* X, size_t ld) {
    if (*flag != 0) return;
    __shared__ double W[AK * AK];
    for (int e = threadIdx.x; e < AK * AK; e += 256) W[e] = out22[AK * AK + e];
    __syncthreads();
    const int idx = blockIdx.x * 256 + threadIdx.x;       // (c, a), a fastest: 32 lanes write 256 contiguous bytes
    const int c = idx >> 5, a = idx & 31;
    if (a >= k) return;
    if (c < n0) {
        L[(size_t)c * ld + n0 + a] = B[(size_t)c * AK + a];
        double s = 0.0;
        for (int b = 0; b <= a; ++b) s = fma(W[a * AK + b], G[(size_t)c * AK + b], s);
        X[(size_t)c * ld + n0 + a] = -s;
    } else if (c < n0 + k) {
        const int b = c - n0;
        L[(size_t)c * ld + n0 + a] = a >= b ? out22[a * AK + b] : 0.0;
        X[(size_t)c * ld + n0 + a] = a >= b ? W[a * AK + b] : 0.0;
    }
}

__global__ void __launch_bounds__(256) dinv_from_x_kernel(const double* __restrict__ X, size_t ld, int tile0, double* Dinv,
                                                          const int* flag) {
    if (flag && *flag != 0) return;
    const int t = tile0 + blockIdx.x;
    const double* src = X + (size_t)t * TB * ld + (size_t)t * TB;
    double* dst = Dinv + (size_t)t * TB * TB;
    for (int idx = threadIdx.x; idx < TB * TB; idx += 256) {
        const int r = idx & (TB - 1), c = idx >> 7;
        dst[idx] = r >= c ? src[(size_t)c * ld + r] : 0.0;
    }
}

// Rows AND columns [r0, r1) of L and X become those of the identity (padding state) for rows < row_end: 1 on the
// diagonal, 0 elsewhere in the lower triangle — and 0 above the diagonal INSIDE the diagonal 128x128 tiles, which the
// tile kernels read as whole tiles (the rest of the upper triangle is never read).
__global__ void __launch_bounds__(256) identity_rows_kernel(double* L, double* X, size_t ld, int r0, int r1, int row_end) {
    const int c = blockIdx.x;                 // column
    const bool c_new = c >= r0 && c < r1;
    const int tile0 = c & ~(TB - 1);          // first row of column c's diagonal tile
    for (int r = min(r0, tile0) + threadIdx.x; r < row_end; r += 256) {
        if (r < tile0) continue;              // above the diagonal tile: never read
        const bool r_new = r >= r0 && r < r1;
        if (!(r_new || c_new)) continue;      // an entry of the old factor
        if (r < c && r >= tile0 + TB) continue;
        const double v = (r == c) ? 1.0 : 0.0;
        L[(size_t)c * ld + r] = v;
        if (X) X[(size_t)c * ld + r] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
size_t append_workspace_doubles(size_t cap) {
    // Pn, B, G (cap x 32 each) + gram partials + S0 (1024) + L22|W (2048) + flag (as one double slot)
    const size_t parts = (cap + GRAM_ROWS - 1) / GRAM_ROWS;
    return 3 * cap * AK + parts * AK * AK + 3 * AK * AK + 8 + ((cap + SPAN - 1) / SPAN) * cap * AK;    // + split-k partials
}

cudaError_t launch_append_slab(const double* xyz, size_t ld, const double* sigma2, int n0, int k, double* L, double* X,
                               double* Dinv, double* ws, size_t cap, const KernParams& kp, int reset_flag,
                               cudaStream_t st) {
    const size_t parts_cap = (cap + GRAM_ROWS - 1) / GRAM_ROWS;
    double* Pn = ws;
    double* B = Pn + cap * AK;
    double* G = B + cap * AK;
    double* part = G + cap * AK;
    double* S0 = part + parts_cap * AK * AK;
    double* out22 = S0 + AK * AK;
    int* flag = reinterpret_cast<int*>(out22 + 2 * AK * AK);
    double* kpart = out22 + 2 * AK * AK + 8;                       // split-k partials of the two skinny products
    if (reset_flag) {
        cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
    }
    const int nblk = (n0 + AK - 1) / AK;
    const int nparts = (n0 + GRAM_ROWS - 1) / GRAM_ROWS;
    append_panel_kernel<<<(n0 * AK + 255) / 256 > 0 ? (n0 * AK + 255) / 256 : 1, 256, 0, st>>>(
        xyz, xyz + ld, xyz + 2 * ld, sigma2, n0, n0, k, Pn, S0, kp);
    const dim3 sgrid((n0 + SROWS - 1) / SROWS, (n0 + SPAN - 1) / SPAN);
    const int fblocks = (n0 * AK + 255) / 256;
    skinny_dmma_kernel<0><<<sgrid, 256, 0, st>>>(X, ld, n0, Pn, kpart);
    skinny_finish_kernel<0><<<fblocks, 256, 0, st>>>(kpart, n0, B);
    append_gram_kernel<<<nparts, 256, 0, st>>>(B, n0, part);
    append_leaf_kernel<<<1, 256, 0, st>>>(S0, part, nparts, out22, flag, n0);
    skinny_dmma_kernel<1><<<sgrid, 256, 0, st>>>(X, ld, n0, B, kpart);
    skinny_finish_kernel<1><<<fblocks, 256, 0, st>>>(kpart, n0, G);
    append_scatter_kernel<<<((n0 + k) * AK + 255) / 256, 256, 0, st>>>(B, G, out22, flag, n0, k, L, X, ld);
    const int t0 = n0 / TB, t1 = (n0 + k - 1) / TB;
    dinv_from_x_kernel<<<t1 - t0 + 1, 256, 0, st>>>(X, ld, t0, Dinv, flag);
    return cudaGetLastError();
}

// Generic launcher of the skinny products (used by the indefinite-tail path, gpr_tail.cu).
cudaError_t launch_skinny(int mode, const double* A, size_t ld, int rows, int kdim, const double* Bm, double* OUT,
                          cudaStream_t st) {
    const int nblk = (rows + AK - 1) / AK;
    if (nblk <= 0) return cudaSuccess;
    if (mode == 0) skinny_tri_kernel<0><<<nblk, 256, 0, st>>>(A, ld, rows, rows, Bm, OUT);
    else if (mode == 1) skinny_tri_kernel<1><<<nblk, 256, 0, st>>>(A, ld, rows, rows, Bm, OUT);
    else skinny_tri_kernel<2><<<nblk, 256, 0, st>>>(A, ld, rows, kdim, Bm, OUT);
    return cudaGetLastError();
}

// Pn[c*32 + a] = k(|p_c - p_{t0+a}|) for c < n0, a < k (and the k x k block S0 of the new points themselves).
cudaError_t launch_append_panel(const double* xyz, size_t ld, const double* sigma2, int n0, int t0, int k, double* Pn,
                                double* S0, const KernParams& kp, cudaStream_t st) {
    const int blocks = (n0 * AK + 255) / 256 > 0 ? (n0 * AK + 255) / 256 : 1;
    append_panel_kernel<<<blocks, 256, 0, st>>>(xyz, xyz + ld, xyz + 2 * ld, sigma2, n0, t0, k, Pn, S0, kp);
    return cudaGetLastError();
}

const int* append_flag_ptr(const double* ws, size_t cap) {
    const size_t parts_cap = (cap + GRAM_ROWS - 1) / GRAM_ROWS;
    return reinterpret_cast<const int*>(ws + 3 * cap * AK + parts_cap * AK * AK + 3 * AK * AK);
}

cudaError_t launch_identity_rows(double* L, double* X, size_t ld, int r0, int r1, int row_end, cudaStream_t st) {
    if (r1 <= r0 || row_end <= r0) return cudaSuccess;
    identity_rows_kernel<<<row_end, 256, 0, st>>>(L, X, ld, r0, r1, row_end);
    return cudaGetLastError();
}

cudaError_t launch_dinv_from_x(const double* X, size_t ld, int tile0, int ntiles, double* Dinv, cudaStream_t st) {
    if (ntiles <= 0) return cudaSuccess;
    dinv_from_x_kernel<<<ntiles, 256, 0, st>>>(X, ld, tile0, Dinv, nullptr);
    return cudaGetLastError();
}

}  // namespace gpr

// gpr_var.cu — K3': predictive variance  v_q = k(0) - || L^-1 k*_q ||^2  for a batch of queries.
//
// Replaces, in the reference's evaluate(gp, query, f, v[, N]) overloads
// (/root/reference/include/gp_regression/gp_regressor.hpp):
//   :255-259, :308-312   Kpq = Kqp^T;  Kqq = k(dist(Q,Q))  (q x q, only its diagonal k(0) is used)
//   :263, :316           V = cholesker.solve(Kpq)           (n x q solve, n^2 flop per query and factor)
//   :265-266, :318-319   V = Kqq - Kqp*V; V.diagonal()      (q x q product for q numbers)
// by a triangular matrix product on the FP64 tensor pipe against the precomputed inverse factor
// X = L^-1 (gpr_factor.cu): V = X * K*^T has no dependency chain, streams X once per query tile through
// L2, and its epilogue reduces the squared column norms so V itself is never written.
//   algorithmic work: n^2 * q flop (one multiply-add per entry of the triangular factor per query).
// Per-tile partial sums are written to a [row tile][query] scratch and summed in a fixed order by a
// second small kernel, so the variance is bit-reproducible (no floating-point atomics) and independent of
// how the queries are batched or sharded.
#include <cstdlib>
#include "gpr_mma.cuh"
#include "gpr_kernels.h"

namespace gpr {

struct VarArgs {
    const double* X; size_t ld; int nb;       // L^-1, lower triangular tiles
    const double* panel; size_t panel_ld;     // K*: element (query, k) at panel[k*panel_ld + query]
    int nqt;                                  // query tiles in this batch (panel_ld / 128)
    double* partial;                          // nb x part_ld
    size_t part_ld;
    int gi, gq, qgroups;                      // row-tile pairs / query tiles per co-scheduled group, number of query groups
};

constexpr int VAR_GI = 8;                     // row-tile pairs per co-scheduled group (GPR_VAR_GI overrides; sweep: tools/var_gi_sweep.sh)

// Task = (pair p, query tile qt): the CTA computes row tile nb-1-p and then row tile p of V = X K*^T for
// its 128 queries.  The k ranges of the two rows add up to nb+1 blocks for EVERY task, so all CTAs of
// the grid run in lockstep from the first wave to the last, and the block order makes the ~148 CTAs
// that are resident together a (VAR_GI pairs) x (gq query tiles) rectangle: each X row tile is streamed
// by gq CTAs at the same time and each K* tile by VAR_GI CTAs at (nearly) the same time, i.e. once from
// HBM and then from L2.  (With one row tile per CTA and a whole row of query tiles resident together,
// every K* tile was fetched from HBM again for each of the nb row tiles: 161 GB of DRAM reads per batch
// for 3.6 GB of operands, ncu profile r1b.  Measured DRAM reads per batch at n = 16384: 4 x 37 -> 45.9 GB,
// 8 x 18 -> 31.2 GB, 16 x 9 -> 35.1 GB, all at the same 35.8 TF/s.)
__global__ void __launch_bounds__(NTHREADS, 1) var_tiles_kernel(VarArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_abort;
    __shared__ double sred[8][32];
    const int per_group = a.gi * a.gq;
    const int g = blockIdx.x / per_group, w = blockIdx.x % per_group;
    const int pg = g / a.qgroups, qg = g % a.qgroups;
    const int p = pg * a.gi + w % a.gi;
    const int qt = qg * a.gq + w / a.gi;
    if (2 * p >= a.nb || qt >= a.nqt) return;          // pairs p < ceil(nb/2)
    if (threadIdx.x == 0) s_abort = 0;
    const TileCoord tc;
    for (int h = 0; h < 2; ++h) {
        const int it = h == 0 ? a.nb - 1 - p : p;
        if (h == 1 && it == a.nb - 1 - p) break;        // odd nb: the middle row tile is its own pair
        Acc acc;
        acc_zero(acc);
        tile_mainloop<STREAM_M, STREAM_M>(acc, a.X + (size_t)it * TB, a.ld, a.panel + (size_t)qt * TB, a.panel_ld,
                                          8 * (it + 1), smem, &s_abort, NoWait());
        // column sums of squares over this tile's 128 rows
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            double s = 0.0;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s = fma(acc[mt][nt][0], acc[mt][nt][0], s);
                s = fma(acc[mt][nt][1], acc[mt][nt][1], s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (tc.t == 0) sred[tc.warp][tc.col<false>(mt) - tc.j0] = s;
        }
        __syncthreads();
        if (threadIdx.x < TB) {
            const int c = threadIdx.x, wj = c >> 5, lc = c & 31;
            a.partial[(size_t)it * a.part_ld + (size_t)qt * TB + c] = sred[2 * wj][lc] + sred[2 * wj + 1][lc];
        }
        __syncthreads();
    }
}

__global__ void var_finalize_kernel(const double* partial, size_t panel_ld, int nb, int q, double k0, double* var) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    double s = 0.0;
    for (int it = 0; it < nb; ++it) s += partial[(size_t)it * panel_ld + i];
    var[i] = k0 - s;
}

cudaError_t launch_var_finalize(const double* partial, size_t panel_ld, int nb, int q, double k0, double* var, cudaStream_t st) {
    if (q <= 0) return cudaSuccess;
    var_finalize_kernel<<<(q + 255) / 256, 256, 0, st>>>(partial, panel_ld, nb, q, k0, var);
    return cudaGetLastError();
}

// panel_pitch: 0, or the real pitch of a panel that is wider than the panel_ld (query tiles) to be processed here.
cudaError_t launch_variance(const double* X, size_t ld, int nb, const double* panel, size_t panel_ld, int q,
                            double* partial, double k0, double* var, cudaStream_t st, size_t panel_pitch) {
    static PerDeviceOnce attr_done;
    const int cur = PerDeviceOnce::current();
    if (!attr_done.done(cur)) {
        cudaError_t e = cudaFuncSetAttribute(var_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TILE_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_done.set(cur);
    }
    if (q <= 0) return cudaSuccess;
    VarArgs a;
    a.X = X; a.ld = ld; a.nb = nb; a.panel = panel; a.panel_ld = panel_pitch ? panel_pitch : panel_ld;
    a.nqt = (int)(panel_ld / TB); a.partial = partial; a.part_ld = panel_ld;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int gi_env = getenv("GPR_VAR_GI") ? atoi(getenv("GPR_VAR_GI")) : 0;
    a.gi = gi_env >= 1 && gi_env <= 64 ? gi_env : VAR_GI;
    a.gq = sms / a.gi > 0 ? sms / a.gi : 1;
    if (a.gq > a.nqt) a.gq = a.nqt;
    a.qgroups = (a.nqt + a.gq - 1) / a.gq;
    const int npairs = (nb + 1) / 2;
    const int pgroups = (npairs + a.gi - 1) / a.gi;
    var_tiles_kernel<<<(unsigned)(pgroups * a.qgroups * a.gi * a.gq), NTHREADS, TILE_SMEM_BYTES, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    var_finalize_kernel<<<(q + 255) / 256, 256, 0, st>>>(partial, panel_ld, nb, q, k0, var);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Variance WITHOUT the explicit inverse: V = L^-1 K*^T by blocked forward substitution, in place in the
// K* panel, one CTA per 128-query tile (the north star's "TRSM for the variance solve v = L^-1 k*";
// reference: cholesker.solve(Kpq), gp_regressor.hpp:263, :316).
//   for row tile i = 0 .. nb-1:   T   = K*_i - sum_{k<i} L_ik V_k      (DMMA mainloop, both operands streamed)
//                                 V_i = Dinv_i T                       (one more DMMA tile product, T resident)
//                                 partial[i][query] = column sums of squares of V_i;  V_i overwrites K*_i
// Same n^2 flop per query as the product with X = L^-1 (nb(nb+1)/2 tile steps per query tile), but no
// one-time n^3/3 inverse and no second n x n matrix per model.  The query tiles are independent, so there
// are no flags and no cross-CTA waits; every CTA walks the same L tiles at (nearly) the same time, so L
// is streamed from HBM once per batch and then served from L2, while the V_k tiles are private to their CTA.
// The panel element (query c, point r) is at panel[r*panel_ld + c]: V_k is read back as the M-major j operand
// exactly like K* in var_tiles_kernel, and written with 16-byte pieces (two adjacent queries per thread).
// ---------------------------------------------------------------------------------------------
struct VarTrsmArgs {
    const double* L; size_t ld; int nb;       // Cholesky factor, lower tiles
    const double* Dinv;                       // nb tiles of 128x128 (ld 128): inverses of the diagonal blocks of L
    double* panel; size_t panel_ld;           // in: K*; out: V = L^-1 K*^T
    int nqt;                                  // query tiles in this batch
    double* partial;                          // nb x panel_ld
};

// smem tile T[s*PM + c] (M-major resident j operand: point s, query c)  =  G - acc,  G = panel tile (coalesced rows)
__device__ __forceinline__ void trsm_residual_to_smem(const Acc& acc, const double* G, size_t ldg, double* T,
                                                      const TileCoord& tc) {
#pragma unroll
    for (int mt = 0; mt < 4; mt += 2) {
        const int c = tc.col<false>(mt);                 // columns c, c+1 belong to m-tiles mt, mt+1
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                double2 v;
                v.x = -GPR_ACC(acc, mt, p, e); v.y = -GPR_ACC(acc, mt + 1, p, e);
                *reinterpret_cast<double2*>(T + (size_t)tc.row(p, e) * PM + c) = v;
            }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < TB * TB / 2; idx += NTHREADS) {
        const int c2 = idx & 63, r = idx >> 6;
        const double2 g = __ldcg(reinterpret_cast<const double2*>(G + (size_t)r * ldg) + c2);
        double2* d = reinterpret_cast<double2*>(T + r * PM) + c2;
        double2 v = *d;
        v.x += g.x; v.y += g.y;
        *d = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NTHREADS, 1) var_trsm_kernel(VarTrsmArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_abort;
    __shared__ double sred[8][32];
    if (threadIdx.x == 0) s_abort = 0;
    const TileCoord tc;
    for (int qt = blockIdx.x; qt < a.nqt; qt += gridDim.x) {
        double* Pq = a.panel + (size_t)qt * TB;
        for (int it = 0; it < a.nb; ++it) {
            Acc acc;
            acc_zero(acc);
            // sum_{k < it} L_{it,k} V_k : i operand = rows of L (M-major), j operand = V (M-major in the panel)
            tile_mainloop<STREAM_M, STREAM_M>(acc, a.L + (size_t)it * TB, a.ld, Pq, a.panel_ld, 8 * it, smem, &s_abort, NoWait());
            double* G = Pq + (size_t)it * TB * a.panel_ld;
            trsm_residual_to_smem(acc, G, a.panel_ld, smem, tc);
            acc_zero(acc);
            // V_it[r][c] = sum_s Dinv_it[r][s] T[s][c]
            tile_mainloop<STREAM_M, RES_M>(acc, a.Dinv + (size_t)it * TB * TB, TB, smem, 0, 8, smem, &s_abort, NoWait());
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                double s = 0.0;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    s = fma(acc[mt][nt][0], acc[mt][nt][0], s);
                    s = fma(acc[mt][nt][1], acc[mt][nt][1], s);
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (tc.t == 0) sred[tc.warp][tc.col<false>(mt) - tc.j0] = s;
            }
            if (it + 1 < a.nb) {
                // V_it replaces K*_it in the panel (read back as the j operand of the later row tiles)
#pragma unroll
                for (int mt = 0; mt < 4; mt += 2) {
                    const int c = tc.col<false>(mt);
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            double2 v;
                            v.x = GPR_ACC(acc, mt, p, e); v.y = GPR_ACC(acc, mt + 1, p, e);
                            *reinterpret_cast<double2*>(G + (size_t)tc.row(p, e) * a.panel_ld + c) = v;
                        }
                }
                __threadfence();
            }
            __syncthreads();
            if (threadIdx.x < TB) {
                const int c = threadIdx.x, wj = c >> 5, lc = c & 31;
                a.partial[(size_t)it * a.panel_ld + (size_t)qt * TB + c] = sred[2 * wj][lc] + sred[2 * wj + 1][lc];
            }
            __syncthreads();
        }
    }
}

cudaError_t launch_variance_trsm(const double* L, size_t ld, int nb, const double* Dinv, double* panel, size_t panel_ld,
                                 int q, double* partial, double k0, double* var, cudaStream_t st) {
    static PerDeviceOnce attr_done;
    const int cur = PerDeviceOnce::current();
    if (!attr_done.done(cur)) {
        cudaError_t e = cudaFuncSetAttribute(var_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TILE_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_done.set(cur);
    }
    if (q <= 0) return cudaSuccess;
    VarTrsmArgs a;
    a.L = L; a.ld = ld; a.nb = nb; a.Dinv = Dinv; a.panel = panel; a.panel_ld = panel_ld;
    a.nqt = (int)(panel_ld / TB); a.partial = partial;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    var_trsm_kernel<<<(unsigned)(a.nqt < sms ? a.nqt : sms), NTHREADS, TILE_SMEM_BYTES, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    var_finalize_kernel<<<(q + 255) / 256, 256, 0, st>>>(partial, panel_ld, nb, q, k0, var);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Small-batch variance (q <= 8): the reference's callers ask for ONE query per call
// (src/gp_node.cpp:1074, include/atlas/atlas_variance.hpp:78,:201).  v = X k* as a matrix-vector
// product, bandwidth-bound: the lower triangle of X is read once for up to 8 queries.
// Thread per row r (coalesced along r), k range split over blockIdx.y; partial dot products go to
// part[split][r][8] and are combined in a fixed order by the finalize kernel.
// ---------------------------------------------------------------------------------------------
constexpr int VG_SPLIT = 8;

template <int Q>
__global__ void __launch_bounds__(256) var_gemv_kernel(const double* __restrict__ X, size_t ld, int N,
                                                       const double* __restrict__ panel, size_t panel_ld,
                                                       double* __restrict__ part) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    const int kspan = N / VG_SPLIT;                    // N is a multiple of 128
    const int k0 = blockIdx.y * kspan;
    const int rmax = min(N - 1, blockIdx.x * 256 + 255);
    const int k1 = min(k0 + kspan, rmax + 1);
    double s[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) s[i] = 0.0;
    if (r < N) {
        for (int k = k0; k < k1; ++k) {
            if (k <= r) {
                const double x = X[(size_t)k * ld + r];
#pragma unroll
                for (int i = 0; i < Q; ++i) s[i] = fma(x, __ldg(panel + (size_t)k * panel_ld + i), s[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) part[((size_t)blockIdx.y * N + r) * 8 + i] = s[i];
    }
}

__global__ void __launch_bounds__(256) var_gemv_finalize_kernel(const double* part, int N, int q, double k0, double* var) {
    __shared__ double red[256];
    for (int i = 0; i < q; ++i) {
        double s = 0.0;
        for (int r = threadIdx.x; r < N; r += 256) {
            double v = 0.0;
            for (int y = 0; y < VG_SPLIT; ++y) v += part[((size_t)y * N + r) * 8 + i];
            s = fma(v, v, s);
        }
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) var[i] = k0 - red[0];
        __syncthreads();
    }
}

// part: VG_SPLIT * N * 8 doubles of scratch.
cudaError_t launch_variance_small(const double* X, size_t ld, int N, const double* panel, size_t panel_ld, int q,
                                  double* part, double k0, double* var, cudaStream_t st) {
    if (q <= 0) return cudaSuccess;
    dim3 grid((N + 255) / 256, VG_SPLIT);
    if (q <= 1) var_gemv_kernel<1><<<grid, 256, 0, st>>>(X, ld, N, panel, panel_ld, part);
    else if (q <= 2) var_gemv_kernel<2><<<grid, 256, 0, st>>>(X, ld, N, panel, panel_ld, part);
    else if (q <= 4) var_gemv_kernel<4><<<grid, 256, 0, st>>>(X, ld, N, panel, panel_ld, part);
    else var_gemv_kernel<8><<<grid, 256, 0, st>>>(X, ld, N, panel, panel_ld, part);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    var_gemv_finalize_kernel<<<1, 256, 0, st>>>(part, N, q, k0, var);
    return cudaGetLastError();
}

}  // namespace gpr

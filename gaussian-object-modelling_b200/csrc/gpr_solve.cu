// gpr_solve.cu — K3: alpha = K^-1 y by two triangular solves with the Cholesky factor, each ONE
// persistent kernel (CTA per 128-row block, dependency flags between blocks).
//
// Replaces gp->alpha = gp->cholesker.solve(gp->Y) in the reference
// (/root/reference/include/gp_regression/gp_regressor.hpp:163, :459).
//   forward  L z = y      : block i accumulates s_i = sum_{k<i} L_ik z_k as the z_k become ready,
//                           then z_i = Dinv_i (y_i - s_i)
//   backward L^T a = z    : block i accumulates s_i = sum_{k>i} L_ki^T a_k, then a_i = Dinv_i^T (z_i - s_i)
// HBM-bound: each solve reads the lower triangle of L once (4 n^2 bytes); Dinv_i = L_ii^-1 comes from
// the factorisation.  Blocks are claimed from an atomic counter in dependency order, so a waiting CTA
// only ever waits on a CTA that is already running (same argument as gpr_factor.cu).
#include "gpr_common.cuh"
#include "gpr_kernels.h"

namespace gpr {

struct TrsvArgs {
    const double* L; size_t ld; int nb;
    const double* Dinv;
    const double* rhs;     // N
    double* out;           // N
    int* ready;            // nb flags
    int* counter;
    int* abort;
};

__global__ void __launch_bounds__(256) trsv_forward_kernel(TrsvArgs a) {
    __shared__ double zs[TB];
    __shared__ double part[2][TB];
    __shared__ int s_blk, s_abort;
    const int tid = threadIdx.x, r = tid & (TB - 1), h = tid >> 7;
    for (;;) {
        if (tid == 0) { s_blk = atomicAdd(a.counter, 1); s_abort = ld_volatile(a.abort) != 0; }
        __syncthreads();
        const int i = s_blk;
        if (i >= a.nb || s_abort) return;
        double s = 0.0;
        for (int k = 0; k < i; ++k) {
            if (tid == 0 && !spin_wait(a.ready + k, a.abort)) s_abort = 1;
            __syncthreads();
            if (s_abort) return;
            if (tid < TB) zs[tid] = __ldcg(a.out + (size_t)k * TB + tid);
            __syncthreads();
            const double* Lp = a.L + ((size_t)k * TB + 64 * h) * a.ld + (size_t)i * TB + r;
#pragma unroll 8
            for (int c = 0; c < 64; ++c) s = fma(Lp[(size_t)c * a.ld], zs[64 * h + c], s);
        }
        part[h][r] = s;
        __syncthreads();
        if (tid < TB) zs[tid] = a.rhs[(size_t)i * TB + tid] - (part[0][tid] + part[1][tid]);
        __syncthreads();
        // z_i = Dinv_i * t   (Dinv_i lower triangular, column-major ld 128)
        const double* D = a.Dinv + (size_t)i * TB * TB + (size_t)(64 * h) * TB + r;
        double z = 0.0;
#pragma unroll 8
        for (int c = 0; c < 64; ++c) z = fma(D[(size_t)c * TB], zs[64 * h + c], z);
        part[h][r] = z;
        __syncthreads();
        if (tid < TB) a.out[(size_t)i * TB + tid] = part[0][tid] + part[1][tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) { __threadfence(); st_release(a.ready + i, 1); }
    }
}

__global__ void __launch_bounds__(256) trsv_backward_kernel(TrsvArgs a) {
    __shared__ double as[TB];
    __shared__ double ts[TB];
    __shared__ int s_blk, s_abort;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        if (tid == 0) { s_blk = atomicAdd(a.counter, 1); s_abort = ld_volatile(a.abort) != 0; }
        __syncthreads();
        if (s_blk >= a.nb || s_abort) return;
        const int i = a.nb - 1 - s_blk;
        // warp w owns columns c = 16w .. 16w+15 of block column i; lane covers rows lane + 32m
        double acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.0;
        for (int k = a.nb - 1; k > i; --k) {
            if (tid == 0 && !spin_wait(a.ready + k, a.abort)) s_abort = 1;
            __syncthreads();
            if (s_abort) return;
            if (tid < TB) as[tid] = __ldcg(a.out + (size_t)k * TB + tid);
            __syncthreads();
            const double* Lp = a.L + ((size_t)i * TB + 16 * warp) * a.ld + (size_t)k * TB + lane;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const double* col = Lp + (size_t)c * a.ld;
                double v = acc[c];
#pragma unroll
                for (int m = 0; m < 4; ++m) v = fma(col[32 * m], as[lane + 32 * m], v);
                acc[c] = v;
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            double v = acc[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) ts[16 * warp + c] = a.rhs[(size_t)i * TB + 16 * warp + c] - v;
        }
        __syncthreads();
        // a_i[c] = sum_r Dinv_i[r][c] t[r]
        const double* D = a.Dinv + (size_t)i * TB * TB + (size_t)(16 * warp) * TB + lane;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const double* col = D + (size_t)c * TB;
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) v = fma(col[32 * m], ts[lane + 32 * m], v);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) a.out[(size_t)i * TB + 16 * warp + c] = v;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) { __threadfence(); st_release(a.ready + i, 1); }
    }
}

// scratch: [0]=counter [2]=abort [4..4+nb) ready.  rhs and out may not alias.
cudaError_t launch_trsv(int backward, const double* L, size_t ld, int nb, const double* Dinv, const double* rhs,
                        double* out, int* scratch, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(int) * (4 + (size_t)nb), st);
    if (e != cudaSuccess) return e;
    TrsvArgs a;
    a.L = L; a.ld = ld; a.nb = nb; a.Dinv = Dinv; a.rhs = rhs; a.out = out;
    a.counter = scratch; a.abort = scratch + 2; a.ready = scratch + 4;
    const int grid = nb < 2 * num_sms ? nb : 2 * num_sms;
    if (backward) trsv_backward_kernel<<<grid, 256, 0, st>>>(a);
    else trsv_forward_kernel<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace gpr

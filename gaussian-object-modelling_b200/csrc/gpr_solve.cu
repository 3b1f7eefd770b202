// gpr_solve.cu — K3: alpha = K^-1 y by two triangular solves with the Cholesky factor, each ONE
// persistent kernel (CTA per 128-row block; a block's result vector carries its own readiness: consumers poll
// the elements they need against a not-ready bit pattern, so no flag round trip and no fence sits in the chain).
//
// Replaces gp->alpha = gp->cholesker.solve(gp->Y) in the reference
// (/root/reference/include/gp_regression/gp_regressor.hpp:163, :459).
//   forward  L z = y      : block i accumulates s_i = sum_{k<i} L_ik z_k as the z_k become ready,
//                           then z_i = Dinv_i (y_i - s_i)
//   backward L^T a = z    : block i accumulates s_i = sum_{k>i} L_ki^T a_k, then a_i = Dinv_i^T (z_i - s_i)
// HBM-bound in volume (each solve reads the lower triangle of L once, 4 n^2 bytes) but latency-bound in
// practice: nb dependent steps.  The step is kept short by taking everything that does not depend on
// the incoming vector off the critical path: Dinv_i = L_ii^-1 (from the factorisation) is staged in
// shared memory with cp.async when the CTA starts, and the next L tile is loaded into registers
// (64 doubles per thread) right after the current one has been consumed, i.e. before the result of the
// next block is awaited.  Blocks are claimed from an atomic counter in dependency order, so a waiting
// CTA only ever waits on a CTA that is already running (same argument as gpr_factor.cu).
#include "gpr_mma.cuh"
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

constexpr size_t TRSV_SMEM = (size_t)TB * TB * sizeof(double);

__device__ __forceinline__ void stage_dinv(double* sD, const double* D) {
    for (int c = threadIdx.x; c < TB * TB / 2; c += 256) cp_async16(sD + 2 * c, D + 2 * c);
    cp_async_commit();
}


// Readiness travels with the data: the launcher fills `out` with an all-ones NaN pattern that no result can have
// (a result with that pattern is stored as the canonical quiet NaN), and a consumer polls the very element it needs.
// This takes one L2 round trip (flag, then data) and both fences out of every step of the dependency chain.
constexpr unsigned long long TRSV_SENTINEL = 0xFFFFFFFFFFFFFFFFull;
__device__ __forceinline__ unsigned long long ld_volatile_u64(const double* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void publish(double* p, double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    if (b == TRSV_SENTINEL) b = 0x7FF8000000000000ull;
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(b) : "memory");
}
// Returns the element once it is there; on abort / timeout (~2 s) raises *s_abort and returns 0.
__device__ __forceinline__ double poll_element(const double* p, int* abort, int* s_abort) {
    unsigned long long b = ld_volatile_u64(p);
    if (b == TRSV_SENTINEL) {
        const long long t0 = clock64();
        for (;;) {
            b = ld_volatile_u64(p);
            if (b != TRSV_SENTINEL) break;
            if (ld_volatile(abort) != 0) { *s_abort = 1; return 0.0; }
            if (clock64() - t0 > 4000000000LL) { atomicExch(abort, 2); *s_abort = 1; return 0.0; }
            __nanosleep(32);
        }
    }
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256, 1) trsv_forward_kernel(TrsvArgs a) {
    extern __shared__ __align__(16) double sD[];   // Dinv_i, column-major ld 128
    __shared__ double zs[TB];
    __shared__ double part[2][TB];
    __shared__ int s_blk, s_abort;
    const int tid = threadIdx.x, r = tid & (TB - 1), h = tid >> 7;
    for (;;) {
        if (tid == 0) { s_blk = atomicAdd(a.counter, 1); s_abort = ld_volatile(a.abort) != 0; }
        __syncthreads();
        const int i = s_blk;
        if (i >= a.nb || s_abort) return;
        stage_dinv(sD, a.Dinv + (size_t)i * TB * TB);
        // thread (r,h) owns row r and the 64 columns [64h, 64h+64) of every tile of block row i
        const double* Lrow = a.L + (size_t)(64 * h) * a.ld + (size_t)i * TB + r;
        double cur[64];
        if (i > 0) {
#pragma unroll
            for (int c = 0; c < 64; ++c) cur[c] = __ldcs(Lrow + (size_t)c * a.ld);
        }
        double s4[4] = {0.0, 0.0, 0.0, 0.0};                   // independent chains: the last step is on the critical path
        for (int k = 0; k < i; ++k) {
            if (tid < TB) zs[tid] = poll_element(a.out + (size_t)k * TB + tid, a.abort, &s_abort);
            __syncthreads();
            if (s_abort) return;
#pragma unroll
            for (int c = 0; c < 64; ++c) s4[c & 3] = fma(cur[c], zs[64 * h + c], s4[c & 3]);
            if (k + 1 < i) {
                const double* nxt = Lrow + (size_t)(k + 1) * TB * a.ld;
#pragma unroll
                for (int c = 0; c < 64; ++c) cur[c] = __ldcs(nxt + (size_t)c * a.ld);
            }
            __syncthreads();                                   // zs is overwritten by the next step's poll
        }
        part[h][r] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        cp_async_wait<0>();
        __syncthreads();
        if (tid < TB) zs[tid] = a.rhs[(size_t)i * TB + tid] - (part[0][tid] + part[1][tid]);
        __syncthreads();
        // z_i = Dinv_i * t
        double z4[4] = {0.0, 0.0, 0.0, 0.0};                   // four independent chains of 16 instead of one of 64
        const double* D = sD + (size_t)(64 * h) * TB + r;
#pragma unroll
        for (int c = 0; c < 64; ++c) z4[c & 3] = fma(D[c * TB], zs[64 * h + c], z4[c & 3]);
        __syncthreads();
        part[h][r] = (z4[0] + z4[1]) + (z4[2] + z4[3]);
        __syncthreads();
        if (tid < TB) publish(a.out + (size_t)i * TB + tid, part[0][tid] + part[1][tid]);
        __syncthreads();                                       // part / zs are reused by the next block of this CTA
    }
}

__global__ void __launch_bounds__(256, 1) trsv_backward_kernel(TrsvArgs a) {
    extern __shared__ __align__(16) double sD[];
    __shared__ double as[TB];
    __shared__ double ts[TB];
    __shared__ int s_blk, s_abort;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        if (tid == 0) { s_blk = atomicAdd(a.counter, 1); s_abort = ld_volatile(a.abort) != 0; }
        __syncthreads();
        if (s_blk >= a.nb || s_abort) return;
        const int i = a.nb - 1 - s_blk;
        stage_dinv(sD, a.Dinv + (size_t)i * TB * TB);
        // warp w owns columns c = 16w .. 16w+15 of block column i; lane covers rows lane + 32m
        const double* Lcol = a.L + ((size_t)i * TB + 16 * warp) * a.ld + lane;
        double cur[16][4];
        double acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.0;
        if (i < a.nb - 1) {
            const double* p = Lcol + (size_t)(a.nb - 1) * TB;
#pragma unroll
            for (int c = 0; c < 16; ++c)
#pragma unroll
                for (int m = 0; m < 4; ++m) cur[c][m] = __ldcs(p + (size_t)c * a.ld + 32 * m);
        }
        for (int k = a.nb - 1; k > i; --k) {
            if (tid < TB) as[tid] = poll_element(a.out + (size_t)k * TB + tid, a.abort, &s_abort);
            __syncthreads();
            if (s_abort) return;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                double v = acc[c];
#pragma unroll
                for (int m = 0; m < 4; ++m) v = fma(cur[c][m], as[lane + 32 * m], v);
                acc[c] = v;
            }
            if (k - 1 > i) {
                const double* p = Lcol + (size_t)(k - 1) * TB;
#pragma unroll
                for (int c = 0; c < 16; ++c)
#pragma unroll
                    for (int m = 0; m < 4; ++m) cur[c][m] = __ldcs(p + (size_t)c * a.ld + 32 * m);
            }
            __syncthreads();                                   // as is overwritten by the next step's poll
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            double v = acc[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) ts[16 * warp + c] = a.rhs[(size_t)i * TB + 16 * warp + c] - v;
        }
        cp_async_wait<0>();
        __syncthreads();
        // a_i[c] = sum_r Dinv_i[r][c] t[r]
        const double* D = sD + (size_t)(16 * warp) * TB + lane;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const double* col = D + (size_t)c * TB;
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) v = fma(col[32 * m], ts[lane + 32 * m], v);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) publish(a.out + (size_t)i * TB + 16 * warp + c, v);
        }
        __syncthreads();                                       // as / ts are reused by the next block of this CTA
    }
}

// scratch: [0]=counter [2]=abort.  rhs and out may not alias; out is filled with the not-ready pattern first.
cudaError_t launch_trsv(int backward, const double* L, size_t ld, int nb, const double* Dinv, const double* rhs,
                        double* out, int* scratch, int num_sms, cudaStream_t st) {
    static PerDeviceOnce attr_done;
    const int dev = PerDeviceOnce::current();
    if (!attr_done.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(trsv_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSV_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(trsv_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSV_SMEM);
        if (e != cudaSuccess) return e;
        attr_done.set(dev);
    }
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(int) * (4 + (size_t)nb), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(out, 0xFF, sizeof(double) * (size_t)nb * TB, st);       // every element "not there yet"
    if (e != cudaSuccess) return e;
    TrsvArgs a;
    a.L = L; a.ld = ld; a.nb = nb; a.Dinv = Dinv; a.rhs = rhs; a.out = out;
    a.counter = scratch; a.abort = scratch + 2; a.ready = scratch + 4;
    const int grid = nb < num_sms ? nb : num_sms;
    if (backward) trsv_backward_kernel<<<grid, 256, TRSV_SMEM, st>>>(a);
    else trsv_forward_kernel<<<grid, 256, TRSV_SMEM, st>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Iterative refinement of alpha (one step after the two triangular solves): r = y - K alpha with K
// re-evaluated entry by entry exactly as cov_build_kernel computed it (the factorisation overwrote K) and
// the products accumulated in double-double (TwoProd by FMA + TwoSum), then L L^T delta = r, alpha += delta.
// The tile Cholesky multiplies by explicit inverses of the 128x128 diagonal blocks (TRSM / TRSV as GEMM /
// GEMV), which costs about one digit of alpha compared with substitution; one refinement step with an
// accurate residual gives it back (tools/accuracy_report.py, profiles/accuracy_*.json).
//   grid (row blocks of 256, chunks of 1024 columns); partial sums [chunk][2][N] reduced in chunk order.
// ---------------------------------------------------------------------------------------------
constexpr int RCH = 1024;

// (s, c) += a * b in double-double: explicit round-to-nearest intrinsics so that nvcc cannot contract the
// product into the following addition (that would destroy the error-free transformations).
__device__ __forceinline__ void dd_add(double& s, double& c, double ph) {
    const double t = __dadd_rn(s, ph);
    const double bb = __dsub_rn(t, s);
    const double e = __dadd_rn(__dsub_rn(s, __dsub_rn(t, bb)), __dsub_rn(ph, bb));      // TwoSum error
    s = t;
    c = __dadd_rn(c, e);
}
__device__ __forceinline__ void dd_add_prod(double& s, double& c, double a, double b) {
    const double ph = __dmul_rn(a, b);
    const double pl = __fma_rn(a, b, -ph);            // exact low part of the product
    dd_add(s, c, ph);
    c = __dadd_rn(c, pl);
}

// 128 rows per CTA: the work items are uniform, so small items keep the last wave over the SMs nearly full
// (n = 8 250: 585 items on 148 SMs instead of 297 = 2 x 148 + 1).
constexpr int RROWS = 128;
__global__ void __launch_bounds__(RROWS) residual_partial_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                                 const double* __restrict__ z, const double* __restrict__ alpha,
                                                                 int n, int N, double* __restrict__ part, KernParams kp) {
    __shared__ double4 sp[RCH];
    const int i = blockIdx.x * RROWS + threadIdx.x;
    const int base = blockIdx.y * RCH;
    for (int k = threadIdx.x; k < RCH; k += RROWS) {
        const int j = base + k;
        sp[k] = j < n ? make_double4(x[j], y[j], z[j], alpha[j]) : make_double4(0.0, 0.0, 0.0, 0.0);
    }
    __syncthreads();
    if (i >= n) return;
    const double xi = x[i], yi = y[i], zi = z[i];
    double s = 0.0, c = 0.0;
    const int lim = min(RCH, n - base);
#pragma unroll 2
    for (int k = 0; k < lim; ++k) {
        const double4 p = sp[k];
        // dist_exact is bit-symmetric in its two points ((-dx)^2 == dx^2), so this is the very entry K_ij that
        // cov_build_kernel wrote and the factorisation read
        dd_add_prod(s, c, kern_value_exact(kp, dist_exact(xi, yi, zi, p.x, p.y, p.z)), p.w);
    }
    part[((size_t)blockIdx.y * 2) * N + i] = s;
    part[((size_t)blockIdx.y * 2 + 1) * N + i] = c;
}

__global__ void __launch_bounds__(256) residual_finish_kernel(const double* __restrict__ part, int nchunks, int n, int N,
                                                              const double* __restrict__ label, const double* __restrict__ sigma2,
                                                              const double* __restrict__ alpha, double* __restrict__ r) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    if (i >= n) { r[i] = 0.0; return; }
    double s = 0.0, c = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) {
        dd_add(s, c, part[((size_t)ch * 2) * N + i]);
        c = __dadd_rn(c, part[((size_t)ch * 2 + 1) * N + i]);
    }
    dd_add_prod(s, c, sigma2[i], alpha[i]);           // the diagonal noise term (gp_regressor.hpp:154-155)
    r[i] = __dsub_rn(__dsub_rn(label[i], s), c);
}

__global__ void __launch_bounds__(256) axpy1_kernel(double* __restrict__ a, const double* __restrict__ d, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) a[i] += d[i];
}

size_t residual_scratch_doubles(int N) { return (size_t)2 * ((N + RCH - 1) / RCH) * N; }

cudaError_t launch_residual(const double* xyz, size_t ld, const double* sigma2, const double* label, const double* alpha,
                            int n, int N, double* part, double* r, const KernParams& kp, cudaStream_t st) {
    const int nchunks = (n + RCH - 1) / RCH;
    dim3 grid((n + RROWS - 1) / RROWS, nchunks);
    residual_partial_kernel<<<grid, RROWS, 0, st>>>(xyz, xyz + ld, xyz + 2 * ld, alpha, n, N, part, kp);
    residual_finish_kernel<<<(N + 255) / 256, 256, 0, st>>>(part, nchunks, n, N, label, sigma2, alpha, r);
    return cudaGetLastError();
}

// dst = src unless the device flag is set (the append's sticky failure flag): lets the host queue the whole append
// without reading the flag in between.
__global__ void __launch_bounds__(256) commit_if_clear_kernel(const int* __restrict__ flag, const double* __restrict__ src,
                                                              double* __restrict__ dst, int n) {
    if (*flag != 0) return;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
cudaError_t launch_commit_if_clear(const int* flag, const double* src, double* dst, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    commit_if_clear_kernel<<<(n + 255) / 256, 256, 0, st>>>(flag, src, dst, n);
    return cudaGetLastError();
}

cudaError_t launch_axpy1(double* a, const double* d, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    axpy1_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, d, n);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// alpha = K^-1 y through the explicit inverse factor X = L^-1 (kept resident for the variance path):
//     v = X y   (lower-triangular matrix-vector product),   alpha = X^T v
// Two dependency-free, bandwidth-bound passes over the triangle of X (4 n^2 bytes each) instead of the two
// flag-chained triangular solves over L (nb dependent steps each).  Used by the incremental append, where X is
// up to date anyway; the refinement step that follows removes the extra rounding of the explicit inverse.
//   lower: CTA = 64 rows x one k-split, thread = (row, k-phase), 16 loads in flight; splits summed in order.
//   upper: CTA = 8 columns, warp = every 8th 32-row chunk of all 8 columns (coalesced, 16 loads in flight), fixed-order
//          shuffle tree and warp sum.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tri_gemv_lower_kernel(const double* __restrict__ X, size_t ld, int n, int kspan,
                                                             const double* __restrict__ in, double* __restrict__ part) {
    __shared__ double ys[256];
    __shared__ double r4[4][64];
    const int tid = threadIdx.x, rl = tid & 63, kq = tid >> 6;
    const int bx = blockIdx.x, by = blockIdx.y;
    const int r = bx * 64 + rl;
    const int kbeg = by * kspan;
    const int kend = min(min(kbeg + kspan, n), bx * 64 + 64);
    double acc = 0.0;
    for (int c0 = kbeg; c0 < kend; c0 += 256) {
        __syncthreads();
        ys[tid] = c0 + tid < kend ? in[c0 + tid] : 0.0;
        __syncthreads();
        const int kl0 = 64 * kq;
        const double* xr = X + (size_t)(c0 + kl0) * ld + r;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            double xv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int c = c0 + kl0 + 16 * b + u;
                xv[u] = (c <= r && c < kend && r < n) ? __ldcs(xr + (size_t)(16 * b + u) * ld) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) acc = fma(xv[u], ys[kl0 + 16 * b + u], acc);
        }
    }
    r4[kq][rl] = acc;
    __syncthreads();
    if (kq == 0 && r < n) part[(size_t)by * n + r] = ((r4[0][rl] + r4[1][rl]) + r4[2][rl]) + r4[3][rl];
}

__global__ void __launch_bounds__(256) tri_gemv_lower_finish_kernel(const double* __restrict__ part, int n, int vs, int kspan,
                                                                    double* __restrict__ out) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= n) return;
    const int nsplit = (min(n, (r / 64) * 64 + 64) + kspan - 1) / kspan;      // splits that reach this row block
    double s = 0.0;
    for (int y = 0; y < nsplit && y < vs; ++y) s += part[(size_t)y * n + r];
    out[r] = s;
}

__global__ void __launch_bounds__(256) tri_gemv_upper_kernel(const double* __restrict__ X, size_t ld, int n,
                                                             const double* __restrict__ v, double* __restrict__ out) {
    __shared__ double red[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 8;                               // columns near 0 are the long ones: they come first
    const double* col = X + (size_t)c0 * ld;
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    // warp w takes the 32-row chunks w, w + 8, ... below the diagonal block; 8 columns x 2 chunks = 16 loads in flight
    const int rbase = (c0 & ~31) + 32 * warp + lane;
    for (int r = rbase; r < n; r += 512) {
        const int r2 = r + 256;
        double x0[8], x1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool cok = c0 + j < n;
            x0[j] = (cok && r >= c0 + j) ? __ldcs(col + (size_t)j * ld + r) : 0.0;
            x1[j] = (cok && r2 < n && r2 >= c0 + j) ? __ldcs(col + (size_t)j * ld + r2) : 0.0;
        }
        const double v0 = v[r], v1 = r2 < n ? v[r2] : 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fma(x1[j], v1, fma(x0[j], v0, acc[j]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double s = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < 8 && c0 + threadIdx.x < n) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        out[c0 + threadIdx.x] = s;
    }
}

// out = X^T (X in) for the leading n x n block of X; scratch: (vs + 1) * n doubles with vs = tri_gemv_splits(n).
int tri_gemv_splits(int n) { int v = (n + 1023) / 1024; return v < 1 ? 1 : (v > 32 ? 32 : v); }

cudaError_t launch_solve_with_inverse(const double* X, size_t ld, int n, const double* in, double* out, double* scratch,
                                      cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int vs = tri_gemv_splits(n);
    const int kspan = ((n + vs - 1) / vs + 255) / 256 * 256;
    double* part = scratch;
    double* v = scratch + (size_t)vs * n;
    dim3 grid((n + 63) / 64, vs);
    tri_gemv_lower_kernel<<<grid, 256, 0, st>>>(X, ld, n, kspan, in, part);
    tri_gemv_lower_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, n, vs, kspan, v);
    tri_gemv_upper_kernel<<<(n + 7) / 8, 256, 0, st>>>(X, ld, n, v, out);
    return cudaGetLastError();
}

}  // namespace gpr

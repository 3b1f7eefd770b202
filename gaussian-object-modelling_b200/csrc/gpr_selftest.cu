// gpr_selftest.cu — device self-tests of the tile engine, reachable through the C-ABI
// (gpr_selftest_*), so that tests/ can check the DMMA fragment maps and the shared-memory leaves
// against numpy one piece at a time.
#include "gpr_mma.cuh"
#include "gpr_leaf.cuh"
#include "gpr_kernels.h"

namespace gpr {

template <bool BK>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_selftest_kernel(const double* A, size_t lda, const double* B,
                                                                    size_t ldb, double* C, size_t ldc, int k) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = 0;
    const int ti = blockIdx.x, tj = blockIdx.y;
    const TileCoord tc;
    Acc acc;
    acc_zero(acc);
    const double* Ag = A + (size_t)ti * TB;
    const double* Bg = BK ? (B + (size_t)tj * TB * ldb) : (B + (size_t)tj * TB);
    tile_mainloop<STREAM_M, BK ? STREAM_K : STREAM_M>(acc, Ag, lda, Bg, ldb, k / KT, smem, &s_abort, NoWait());
    store_tile<BK, 1>(acc, C + (size_t)tj * TB * ldc + (size_t)ti * TB, ldc, tc);
}

cudaError_t launch_gemm_selftest(const double* A, size_t lda, const double* B, size_t ldb, int b_kmajor, double* C,
                                 size_t ldc, int mt, int nt, int k, cudaStream_t st) {
    cudaError_t e;
    dim3 grid(mt, nt);
    if (b_kmajor) {
        e = cudaFuncSetAttribute(gemm_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        gemm_selftest_kernel<true><<<grid, NTHREADS, TILE_SMEM_BYTES, st>>>(A, lda, B, ldb, C, ldc, k);
    } else {
        e = cudaFuncSetAttribute(gemm_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        gemm_selftest_kernel<false><<<grid, NTHREADS, TILE_SMEM_BYTES, st>>>(A, lda, B, ldb, C, ldc, k);
    }
    return cudaGetLastError();
}

__global__ void __launch_bounds__(NTHREADS, 1) leaf_selftest_kernel(double* tile, double* inv, int* info, long long* cycles) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_fail;
    __shared__ double s_inv[TB];
    if (threadIdx.x == 0) s_fail = 1 << 20;
    for (int idx = threadIdx.x; idx < TB * TB; idx += NTHREADS) smem[(idx >> 7) * PM + (idx & 127)] = tile[idx];
    __syncthreads();
    long long c0 = clock64();
    potrf128_smem(smem, s_inv, &s_fail, cycles ? cycles + 2 : nullptr);
    long long c1 = clock64();
    if (threadIdx.x == 0) *info = s_fail < TB ? s_fail + 1 : 0;
    store_lower_tile(smem, tile, TB);
    __syncthreads();
    long long c2 = clock64();
    trinv128_smem(smem, s_inv, smem + R0_DBL);
    long long c3 = clock64();
    store_lower_tile(smem, inv, TB);
    if (cycles && threadIdx.x == 0) { cycles[0] = c1 - c0; cycles[1] = c3 - c2; }
}

cudaError_t launch_leaf_selftest(double* tile, double* inv, int* info, cudaStream_t st, long long* cycles) {
    cudaError_t e = cudaFuncSetAttribute(leaf_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    leaf_selftest_kernel<<<1, NTHREADS, TILE_SMEM_BYTES, st>>>(tile, inv, info, cycles);
    return cudaGetLastError();
}

}  // namespace gpr

// ---------------------------------------------------------------------------------------------
// Peak probes: raw issue rate of the FP64 tensor pipe (DMMA.8x8x4) and of the FP64 FMA pipe, measured
// with CUDA events.  These are the denominators for the Cholesky / variance rooflines when
// MEASURED_PEAKS.json has no FP64 entry.
// ---------------------------------------------------------------------------------------------
namespace gpr {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

// Both pipes at once: per iteration 16 DMMA (128 FMA per thread) interleaved with 16*REP independent DFMA
// per thread.  Shows whether DMMA and DFMA are served by the same FP64 datapath (rates do not add) or not.
template <int REP>
__global__ void __launch_bounds__(256) dmix_peak_kernel(double* out, int iters) {
    double c[16][2], f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; f[i] = i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int r = 0; r < REP; ++r) f[(i * REP + r) & 15] = fma(f[(i * REP + r) & 15], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1] + f[i];
    if (s == 123.456) out[0] = s;
}

// which: 0 DMMA, 1 DFMA, 2 / 3: 16 DMMA (128 FMA per thread) mixed with 32 / 128 DFMA per thread.  Returns achieved TFLOP/s (2 flop per multiply-add) in *tflops.
cudaError_t run_peak_probe(int which, int ctas_per_sm, double* tflops) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d;
    cudaError_t e = cudaMalloc((void**)&d, 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = which == 1 ? 20000 : 4000;
    const int grid = sms * ctas_per_sm;
    float best = 1e30f;
    for (int rep = 0; rep < 8; ++rep) {                  // rep 0 warms up; best of the rest
        cudaEventRecord(a);
        if (which == 0) dmma_peak_kernel<<<grid, 256>>>(d, iters);
        else if (which == 1) dfma_peak_kernel<<<grid, 256>>>(d, iters);
        else if (which == 2) dmix_peak_kernel<2><<<grid, 256>>>(d, iters);
        else dmix_peak_kernel<8><<<grid, 256>>>(d, iters);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) return e;
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    const double dmma_fma = 16.0 * 8 * 8 * 4 / 32.0;                              // DMMA: 256 FMA per warp instr
    const double per_thread_fma = which == 0 ? dmma_fma : which == 1 ? 16.0 : which == 2 ? dmma_fma + 32.0 : dmma_fma + 128.0;
    const double flops = 2.0 * per_thread_fma * iters * 256.0 * grid;
    *tflops = flops / (best * 1e-3) / 1e12;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    return cudaGetLastError();
}

}  // namespace gpr

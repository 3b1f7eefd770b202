// gpr_common.cuh — shared definitions for the sm_100a GP-regression kernels.
//
// Storage conventions:
//   * every dense matrix is column-major FP64 with a leading dimension that is a multiple of 128;
//   * the training set of n points is padded to N = 128*ceil(n/128); padded rows/columns of K are
//     the identity, padded labels/alpha are 0, so every kernel works on whole 128x128 tiles;
//   * only the lower triangle (tile row >= tile column) of K / L / L^-1 is ever read.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <atomic>

namespace gpr {

// cudaFuncSetAttribute applies to the CURRENT device only: a launcher that raises a kernel's dynamic shared-memory
// limit must do so once per device, not once per process (a context may sit on any device, and the query shards of
// one context run on several).
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    static int current() { int d = 0; cudaGetDevice(&d); return d; }
    bool done(int dev) const { return dev >= 0 && dev < 64 && ((mask.load(std::memory_order_acquire) >> dev) & 1ull); }
    void set(int dev) { if (dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release); }
};

constexpr int TB = 128;          // tile edge
constexpr int NTHREADS = 256;    // threads per CTA in all tile kernels

// Covariance functor parameters, precomputed on the host exactly as the reference's constructors do
// (kernels/thin_plate.hpp:28-32, kernels/gaussian.hpp:36-42, kernels/laplace.hpp:58-63).
struct KernParams {
    int kind;        // 0 ThinPlate, 1 Gaussian, 2 Laplace
    double p0, p1;   // R | sigma,length | sigma,length
    double R3;       // R^3
    double amp;      // sigma^2 (Gaussian) or 2*sigma (Laplace)
    double inv;      // 1/length^2 (Gaussian) or 1/length (Laplace)
};

// k(d), operation order of the reference (kernels/thin_plate.hpp:14, gaussian.hpp:17-18,
// laplace.hpp:39-40) with explicit round-to-nearest multiplies/adds and NO fma contraction, so that
// the thin-plate covariance is bit-identical to the CPU oracle (which is built -ffp-contract=off).
__device__ __forceinline__ double kern_value_exact(const KernParams& kp, double d) {
    if (kp.kind == 0) {
        double a = __dmul_rn(__dmul_rn(__dmul_rn(2.0, d), d), d);
        double b = __dmul_rn(__dmul_rn(__dmul_rn(3.0, kp.p0), d), d);
        return __dadd_rn(__dsub_rn(a, b), kp.R3);
    }
    return __dmul_rn(kp.amp, exp(__dmul_rn(__dmul_rn(-1.0, d), kp.inv)));
}

// Euclidean distance in difference form, sqrt(dx^2+dy^2+dz^2) accumulated left to right without fma
// (documented deviation (i) from gp_regressor.hpp:548-557).
__device__ __forceinline__ double dist_exact(double ax, double ay, double az, double bx, double by, double bz) {
    double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return sqrt(s);
}

// Fast forms for the prediction kernels (fma contraction allowed; same formulas).
template <int KIND>
__device__ __forceinline__ double kern_value(const KernParams& kp, double d) {
    if (KIND == 0) {
        // 2d^3 - 3Rd^2 + R^3 = d^2 (2d - 3R) + R^3
        return fma(d * d, fma(2.0, d, -3.0 * kp.p0), kp.R3);
    }
    return kp.amp * exp(-d * kp.inv);
}
// The reference's computediff: -6(R-d) for ThinPlate (thin_plate.hpp:19), -(inv)*k(d) otherwise
// (gaussian.hpp:24-25, laplace.hpp:46-47).  kval is k(d) when already available.
template <int KIND>
__device__ __forceinline__ double kern_diff(const KernParams& kp, double d, double kval) {
    if (KIND == 0) return -6.0 * (kp.p0 - d);
    return -kp.inv * kval;
}

// ---- small PTX helpers -----------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spin until *flag != 0.  Returns false if another CTA raised *abort or ~2 s passed (then raises it).
// Tasks are claimed from an atomic counter in an order in which every dependency has a smaller
// index, so a waiting CTA always waits on a CTA that is already running: no deadlock by construction;
// the time limit only guards against bugs.
__device__ __forceinline__ bool spin_wait(const int* flag, int* abort) {
    if (ld_acquire(flag) != 0) return true;
    long long t0 = clock64();
    for (;;) {
        if (ld_acquire(flag) != 0) return true;
        if (ld_volatile(abort) != 0) return false;
        if (clock64() - t0 > 4000000000LL) { atomicExch(abort, 2); return false; }
        __nanosleep(64);
    }
}

// Block-uniform length of the leading run of ready items among [0, n): every thread tests a strided
// subset with flag_ok(t) (an acquire load), the first failures are min-reduced through *s_min, which the
// caller has set to a value >= n before the preceding barrier.  Contains one __syncthreads().
template <class FlagOk>
__device__ __forceinline__ int ready_prefix(int n, FlagOk flag_ok, int* s_min) {
    int first_bad = n;
    for (int t = threadIdx.x; t < n; t += NTHREADS) {
        if (!flag_ok(t)) { first_bad = t; break; }
    }
    first_bad = __reduce_min_sync(0xffffffffu, first_bad);
    if ((threadIdx.x & 31) == 0) atomicMin(s_min, first_bad);
    __syncthreads();
    return *s_min;
}

}  // namespace gpr

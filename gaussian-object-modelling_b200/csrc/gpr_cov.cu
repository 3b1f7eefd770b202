// gpr_cov.cu — K1: fused covariance build.  One pass: pairwise distance -> kernel -> (+sigma^2 on
// the diagonal) -> K, each entry written exactly once; max pairwise distance (Model::R) reduced in
// the same pass.
//
// Replaces, in the reference's create<>() (/root/reference/include/gp_regression/gp_regressor.hpp):
//   :132  buildEuclideanDistanceMatrix (GEMM + sqrt, n^2 doubles written)       -> fused, never stored
//   :135  R = Kpp.maxCoeff()                                                    -> atomicMax in-kernel
//   :144-159  serial double loop applying kernel_->compute and adding sigma2    -> epilogue
// and, in update<>() :397-440, the same for the appended rows.
//
// Only the tiles on or below the diagonal are produced (the factorisation never reads the rest):
// algorithmic bytes = 8 * 128^2 * nb(nb+1)/2 written + 24n read.
// Layout: training coordinates are SoA (as gp_regression::Data, gp_regressor.hpp:49-55), staged per
// tile in shared memory; each thread owns two adjacent rows (one 16-byte store) and walks 32 columns,
// so a warp writes 512 contiguous bytes per column.
// Distances use the difference form and the kernel is evaluated without fma contraction so that the
// thin-plate K is bit-identical to the CPU oracle (SURVEY F8, §8c deviation (i)).
#include "gpr_common.cuh"
#include "gpr_kernels.h"

namespace gpr {

struct CovArgs {
    const double* x; const double* y; const double* z;   // padded to N (padding coordinates are 0)
    const double* sigma2;                                // N entries (0 where absent)
    int n;                                               // real points
    int nb;                                              // N / 128
    int tile_row0;                                       // first tile row to build (append: rows >= this)
    double* K; size_t ld;
    unsigned long long* rmax_bits;                       // max distance as raw bits (distances are >= 0)
    KernParams kp;
};

__global__ void __launch_bounds__(NTHREADS) cov_build_kernel(CovArgs a) {
    // linear block -> lower-triangular tile (ti >= tj), restricted to ti >= tile_row0
    int b = blockIdx.x;
    int ti = a.tile_row0;
    while (b >= ti + 1) { b -= ti + 1; ++ti; }
    const int tj = b;

    __shared__ double sx[TB], sy[TB], sz[TB];
    const int tid = threadIdx.x;
    if (tid < TB) {
        sx[tid] = a.x[tj * TB + tid]; sy[tid] = a.y[tj * TB + tid]; sz[tid] = a.z[tj * TB + tid];
    }
    const int r2 = tid & 63;            // row pair within the tile
    const int cq = tid >> 6;            // column phase 0..3
    const int gi = ti * TB + 2 * r2;    // global row of the first of my two rows
    const double x0 = a.x[gi], y0 = a.y[gi], z0 = a.z[gi];
    const double x1 = a.x[gi + 1], y1 = a.y[gi + 1], z1 = a.z[gi + 1];
    const double s0 = a.sigma2[gi], s1 = a.sigma2[gi + 1];
    __syncthreads();

    double dmax = 0.0;
#pragma unroll 4
    for (int c = 0; c < 32; ++c) {
        const int cl = cq + 4 * c;
        const int gj = tj * TB + cl;
        const double d0 = dist_exact(x0, y0, z0, sx[cl], sy[cl], sz[cl]);
        const double d1 = dist_exact(x1, y1, z1, sx[cl], sy[cl], sz[cl]);
        double2 v;
        v.x = kern_value_exact(a.kp, d0);
        v.y = kern_value_exact(a.kp, d1);
        const bool colreal = gj < a.n;
        if (gi == gj) v.x = __dadd_rn(v.x, s0);            // gp_regressor.hpp:154-155
        if (gi + 1 == gj) v.y = __dadd_rn(v.y, s1);
        if (!(colreal && gi < a.n)) v.x = (gi == gj) ? 1.0 : 0.0;          // identity padding
        else dmax = fmax(dmax, d0);
        if (!(colreal && gi + 1 < a.n)) v.y = (gi + 1 == gj) ? 1.0 : 0.0;
        else dmax = fmax(dmax, d1);
        *reinterpret_cast<double2*>(a.K + (size_t)gj * a.ld + gi) = v;
    }
    // max distance: warp shuffle reduce, one atomic per warp (order-independent, hence deterministic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    if ((tid & 31) == 0) atomicMax(a.rmax_bits, (unsigned long long)__double_as_longlong(dmax));
}

cudaError_t launch_cov_build(const double* x, const double* y, const double* z, const double* sigma2, int n, int nb,
                             int tile_row0, double* K, size_t ld, unsigned long long* rmax_bits, const KernParams& kp,
                             cudaStream_t st) {
    CovArgs a;
    a.x = x; a.y = y; a.z = z; a.sigma2 = sigma2; a.n = n; a.nb = nb; a.tile_row0 = tile_row0;
    a.K = K; a.ld = ld; a.rmax_bits = rmax_bits; a.kp = kp;
    long long tiles = (long long)nb * (nb + 1) / 2 - (long long)tile_row0 * (tile_row0 + 1) / 2;
    if (tiles <= 0) return cudaSuccess;
    cov_build_kernel<<<(unsigned)tiles, NTHREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace gpr

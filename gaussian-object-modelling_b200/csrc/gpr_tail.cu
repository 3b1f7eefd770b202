// gpr_tail.cu — indefinite covariance matrices: block L D L^T with ONE dense trailing pivot block.
//
// The reference factorises K with Eigen's diagonally pivoted LDLT
// (/root/reference/include/gp_regression/gp_regressor.hpp:81, :161-163), which also works when K is
// indefinite — and in the ROS node's real configuration it is: ThinPlate(2.0) with 15 external points on
// the r = 2 sphere (src/gp_node.cpp:16, :821-849, :919) gives pair distances up to 3.9 > R and three
// negative eigenvalues.  A plain Cholesky stops at the first external point.
//
// Here: when the tile Cholesky meets a non-positive pivot at point p and at most 256 points remain (the
// external points are appended last, src/gp_node.cpp:898-914), the leading block is kept as it is —
//     K = [[A, P], [P^T, C]],   A = L L^T  (p x p, SPD),   X = L^-1
// — and the trailing m = n - p points are eliminated as one dense block:
//     B = X P (p x m),   S = C - B^T B  (m x m, symmetric, indefinite),   Z = X^T B = A^-1 P
//     alpha_2 = S^-1 (y_2 - B^T X y_1),   alpha_1 = A^-1 y_1 - Z alpha_2
//     var(q)  = k(0) - |X k_1|^2 - w^T S^-1 w,   w = k_2 - Z^T k_1
// which is K^-1 exactly (a block L D L^T whose D has one m x m block), i.e. what the reference's pivoted
// LDLT computes, to cond(K) eps.  S^-1 (m <= 256) is formed on the host by Gaussian elimination with
// partial pivoting.  B and Z are stored in slabs of 32 columns: slab s, row r, column a at
// [(s*ldr + r)*32 + a] — the layout the skinny products of gpr_append.cu produce.
#include "gpr_common.cuh"
#include "gpr_kernels.h"

namespace gpr {

constexpr int TK = 32;          // slab width (== AK of gpr_append.cu)

// C[a*mp + b] = k(|t_a - t_b|) + [a==b] sigma2_a for the m tail points at [p, p+m); identity outside.
__global__ void __launch_bounds__(256) tail_cc_kernel(const double* x, const double* y, const double* z,
                                                      const double* sigma2, int p, int m, int mp, double* C,
                                                      KernParams kp) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= mp * mp) return;
    const int a = e / mp, b = e % mp;
    double v = (a == b) ? 1.0 : 0.0;
    if (a < m && b < m) {
        v = kern_value_exact(kp, dist_exact(x[p + a], y[p + a], z[p + a], x[p + b], y[p + b], z[p + b]));
        if (a == b) v = __dadd_rn(v, sigma2[p + a]);
    }
    C[e] = v;
}

// Partial Gram blocks of B: part[(pair*nparts + blk)*1024 + a*32 + b] = sum_{r in 256-row block} Bsa[r][a] Bsb[r][b]
// for the slab pair (sa >= sb) number `pair` = sa(sa+1)/2 + sb.
__global__ void __launch_bounds__(256) tail_gram_kernel(const double* __restrict__ B, size_t ldr, int p, int nparts,
                                                        double* __restrict__ part) {
    __shared__ double Sa[64 * TK], Sb[64 * TK];
    const int pair = blockIdx.y;
    int sa = 0;
    while ((sa + 1) * (sa + 2) / 2 <= pair) ++sa;
    const int sb = pair - sa * (sa + 1) / 2;
    const int tid = threadIdx.x, a = tid & 31, b0 = (tid >> 5) * 4;
    const int r0 = blockIdx.x * 256;
    const double* Ba = B + (size_t)sa * ldr * TK;
    const double* Bb = B + (size_t)sb * ldr * TK;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int rc = 0; rc < 256; rc += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * TK; e += 256) {
            const int r = r0 + rc + (e >> 5);
            Sa[e] = r < p ? Ba[(size_t)r * TK + (e & 31)] : 0.0;
            Sb[e] = r < p ? Bb[(size_t)r * TK + (e & 31)] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 64; ++r) {
            const double va = Sa[r * TK + a];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fma(va, Sb[r * TK + b0 + j], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) part[((size_t)pair * nparts + blockIdx.x) * TK * TK + a * TK + b0 + j] = acc[j];
}

// S = C - B^T B (both triangles), partial sums added in a fixed order.  One CTA per slab pair.
__global__ void __launch_bounds__(256) tail_schur_kernel(const double* __restrict__ C, const double* __restrict__ part,
                                                         int nparts, int mp, double* __restrict__ S) {
    const int pair = blockIdx.x;
    int sa = 0;
    while ((sa + 1) * (sa + 2) / 2 <= pair) ++sa;
    const int sb = pair - sa * (sa + 1) / 2;
    for (int e = threadIdx.x; e < TK * TK; e += 256) {
        const int a = e >> 5, b = e & 31;
        double s = 0.0;
        for (int blk = 0; blk < nparts; ++blk) s += part[((size_t)pair * nparts + blk) * TK * TK + e];
        const int ga = sa * TK + a, gb = sb * TK + b;
        const double v = C[(size_t)ga * mp + gb] - s;
        S[(size_t)ga * mp + gb] = v;
        S[(size_t)gb * mp + ga] = v;
    }
}

// t[a] = y2[a] - sum_{r<p} B[r][a] zf[r]     (one CTA per slab, 8 row-phases x 32 columns, fixed-order tree)
__global__ void __launch_bounds__(256) tail_t_kernel(const double* __restrict__ B, size_t ldr, int p, int m,
                                                     const double* __restrict__ zf, const double* __restrict__ label,
                                                     double* __restrict__ t) {
    __shared__ double red[8][TK];
    const int s = blockIdx.x, a = threadIdx.x & 31, ph = threadIdx.x >> 5;
    const double* Bs = B + (size_t)s * ldr * TK;
    double acc = 0.0;
    for (int r = ph; r < p; r += 8) acc = fma(Bs[(size_t)r * TK + a], zf[r], acc);
    red[ph][a] = acc;
    __syncthreads();
    if (ph == 0) {
        double v = 0.0;
        for (int k = 0; k < 8; ++k) v += red[k][a];
        const int ga = s * TK + a;
        t[ga] = ga < m ? label[p + ga] - v : 0.0;
    }
}

// alpha_2 = Sinv t (one CTA), written to alpha[p + a] and to a2[a].
__global__ void __launch_bounds__(256) tail_alpha2_kernel(const double* __restrict__ Sinv, int mp, int m,
                                                          const double* __restrict__ t, double* __restrict__ a2,
                                                          double* __restrict__ alpha, int p) {
    for (int a = threadIdx.x; a < mp; a += 256) {
        double v = 0.0;
        if (a < m) for (int b = 0; b < m; ++b) v = fma(Sinv[(size_t)a * mp + b], t[b], v);
        a2[a] = v;
        if (a < m) alpha[p + a] = v;
    }
}

// alpha_1[r] = z1[r] - sum_a Z[r][a] alpha_2[a]   (in place on alpha[0..p))
__global__ void __launch_bounds__(256) tail_alpha1_kernel(const double* __restrict__ Z, size_t ldr, int p, int nslab,
                                                          const double* __restrict__ a2, double* __restrict__ alpha) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= p) return;
    double v = 0.0;
    for (int s = 0; s < nslab; ++s) {
        const double* zr = Z + ((size_t)s * ldr + r) * TK;
#pragma unroll 8
        for (int a = 0; a < TK; ++a) v = fma(zr[a], a2[s * TK + a], v);
    }
    alpha[r] -= v;
}

// var[q] -= w^T Sinv w with w[a] = k(|q - t_a|) - W[q][a], W = Z^T k_1 from the skinny product (slab layout,
// row = query, ldq rows per slab).  Thread per query; Sinv and the tail points are staged in shared memory
// by 32 x 32 blocks.
template <int KIND>
__global__ void __launch_bounds__(256) tail_var_kernel(const double* __restrict__ qx, const double* __restrict__ qy,
                                                       const double* __restrict__ qz, int q, const double* x,
                                                       const double* y, const double* z, int p, int m, int mp,
                                                       const double* __restrict__ W, size_t ldq,
                                                       const double* __restrict__ Sinv, double* __restrict__ var,
                                                       KernParams kp) {
    extern __shared__ double sh[];                 // mp (w scratch is per thread in local arrays: m <= 256 -> loop by slabs)
    double* sS = sh;                               // 32 x 32 block of Sinv
    double* sT = sh + TK * TK;                     // tail coordinates: 3 x mp
    const int qi = blockIdx.x * 256 + threadIdx.x;
    const bool real = qi < q;
    for (int e = threadIdx.x; e < 3 * mp; e += 256) {
        const int c = e / mp, a = e % mp;
        const double* src = c == 0 ? x : (c == 1 ? y : z);
        sT[e] = a < m ? src[p + a] : 0.0;
    }
    __syncthreads();
    const double X0 = real ? qx[qi] : 0.0, Y0 = real ? qy[qi] : 0.0, Z0 = real ? qz[qi] : 0.0;
    const int nslab = mp / TK;
    double corr = 0.0;
    // corr = sum_{sa} sum_{sb} w_sa^T Sinv[sa][sb] w_sb, w recomputed per slab (32 values in registers)
    for (int sa = 0; sa < nslab; ++sa) {
        double wa[TK];
#pragma unroll
        for (int a = 0; a < TK; ++a) {
            const int ga = sa * TK + a;
            const double dx = X0 - sT[ga], dy = Y0 - sT[mp + ga], dz = Z0 - sT[2 * mp + ga];
            const double d = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
            const double kv = ga < m ? kern_value<KIND>(kp, d) : 0.0;
            wa[a] = real ? kv - W[((size_t)sa * ldq + qi) * TK + a] : 0.0;
        }
        for (int sb = 0; sb < nslab; ++sb) {
            __syncthreads();
            for (int e = threadIdx.x; e < TK * TK; e += 256)
                sS[e] = Sinv[(size_t)(sa * TK + (e >> 5)) * mp + sb * TK + (e & 31)];
            __syncthreads();
            double wb[TK];
#pragma unroll
            for (int b = 0; b < TK; ++b) {
                const int gb = sb * TK + b;
                const double dx = X0 - sT[gb], dy = Y0 - sT[mp + gb], dz = Z0 - sT[2 * mp + gb];
                const double d = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
                const double kv = gb < m ? kern_value<KIND>(kp, d) : 0.0;
                wb[b] = real ? kv - W[((size_t)sb * ldq + qi) * TK + b] : 0.0;
            }
#pragma unroll 4
            for (int a = 0; a < TK; ++a) {
                double row = 0.0;
#pragma unroll
                for (int b = 0; b < TK; ++b) row = fma(sS[a * TK + b], wb[b], row);
                corr = fma(wa[a], row, corr);
            }
        }
    }
    if (real) var[qi] -= corr;
}

// counts[i] = #{ j != i : K_ij^2 > K_ii K_jj }: the number of points whose 2x2 minor with point i is indefinite —
// no positive definite block can hold both.  For the thin-plate kernel these are the points farther than ~R
// from i; an outlier conflicts with most of the set.  Used by the host to choose which points to move into the
// trailing pivot block (gpr_c_api.cu).  Thread per point, the others staged in shared memory.
__global__ void __launch_bounds__(256) conflict_count_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                             const double* __restrict__ z, const double* __restrict__ sigma2,
                                                             int n, int* __restrict__ counts, KernParams kp) {
    __shared__ double4 sp[256];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool real = i < n;
    const double xi = real ? x[i] : 0.0, yi = real ? y[i] : 0.0, zi = real ? z[i] : 0.0;
    const double k0 = kp.kind == 0 ? kp.R3 : kp.amp;
    const double kii = k0 + (real ? sigma2[i] : 0.0);
    int cnt = 0;
    for (int base = 0; base < n; base += 256) {
        __syncthreads();
        const int j = base + threadIdx.x;
        sp[threadIdx.x] = j < n ? make_double4(x[j], y[j], z[j], k0 + sigma2[j]) : make_double4(0.0, 0.0, 0.0, 0.0);
        __syncthreads();
        const int lim = min(256, n - base);
        for (int k = 0; k < lim; ++k) {
            if (base + k == i) continue;
            const double4 q = sp[k];
            const double kij = kern_value_exact(kp, dist_exact(xi, yi, zi, q.x, q.y, q.z));
            if (kij * kij > kii * q.w) ++cnt;
        }
    }
    if (real) counts[i] = cnt;
}

cudaError_t launch_conflict_counts(const double* xyz, size_t ld, const double* sigma2, int n, int* counts,
                                   const KernParams& kp, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    conflict_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(xyz, xyz + ld, xyz + 2 * ld, sigma2, n, counts, kp);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_tail_cc(const double* xyz, size_t ld, const double* sigma2, int p, int m, int mp, double* C,
                           const KernParams& kp, cudaStream_t st) {
    tail_cc_kernel<<<(mp * mp + 255) / 256, 256, 0, st>>>(xyz, xyz + ld, xyz + 2 * ld, sigma2, p, m, mp, C, kp);
    return cudaGetLastError();
}

size_t tail_gram_part_doubles(int p, int mp) {
    const int ns = mp / TK;
    return (size_t)(ns * (ns + 1) / 2) * ((p + 255) / 256) * TK * TK;
}

cudaError_t launch_tail_schur(const double* B, size_t ldr, int p, int mp, const double* C, double* part, double* S,
                              cudaStream_t st) {
    const int ns = mp / TK, npairs = ns * (ns + 1) / 2, nparts = (p + 255) / 256;
    dim3 grid(nparts, npairs);
    tail_gram_kernel<<<grid, 256, 0, st>>>(B, ldr, p, nparts, part);
    tail_schur_kernel<<<npairs, 256, 0, st>>>(C, part, nparts, mp, S);
    return cudaGetLastError();
}

cudaError_t launch_tail_alpha(const double* B, const double* Z, size_t ldr, int p, int m, int mp, const double* zf,
                              const double* label, const double* Sinv, double* t, double* a2, double* alpha,
                              cudaStream_t st) {
    const int ns = mp / TK;
    tail_t_kernel<<<ns, 256, 0, st>>>(B, ldr, p, m, zf, label, t);
    tail_alpha2_kernel<<<1, 256, 0, st>>>(Sinv, mp, m, t, a2, alpha, p);
    tail_alpha1_kernel<<<(p + 255) / 256, 256, 0, st>>>(Z, ldr, p, ns, a2, alpha);
    return cudaGetLastError();
}

cudaError_t launch_tail_var(const double* qx, const double* qy, const double* qz, int q, const double* xyz, size_t ld,
                            int p, int m, int mp, const double* W, size_t ldq, const double* Sinv, double* var,
                            const KernParams& kp, cudaStream_t st) {
    if (q <= 0) return cudaSuccess;
    const size_t sh = (size_t)(TK * TK + 3 * mp) * sizeof(double);
    const int grid = (q + 255) / 256;
    switch (kp.kind) {
        case 0: tail_var_kernel<0><<<grid, 256, sh, st>>>(qx, qy, qz, q, xyz, xyz + ld, xyz + 2 * ld, p, m, mp, W, ldq, Sinv, var, kp); break;
        case 1: tail_var_kernel<1><<<grid, 256, sh, st>>>(qx, qy, qz, q, xyz, xyz + ld, xyz + 2 * ld, p, m, mp, W, ldq, Sinv, var, kp); break;
        default: tail_var_kernel<2><<<grid, 256, sh, st>>>(qx, qy, qz, q, xyz, xyz + ld, xyz + 2 * ld, p, m, mp, W, ldq, Sinv, var, kp); break;
    }
    return cudaGetLastError();
}

}  // namespace gpr

// gpr_predict.cu — K4: fused prediction.  For every query the cross-covariance row k(|q - p_j|) is
// built on the fly and reduced against alpha in registers; K(X*,X) is never materialised for the
// mean / gradient.  When the variance is requested the same pass also writes the cross-covariance
// panel of the current query batch (the right-hand side of the variance solve, gpr_var.cu).
//
// Replaces, in the reference's evaluate() overloads
// (/root/reference/include/gp_regression/gp_regressor.hpp):
//   :240, :300, :348   buildEuclideanDistanceMatrix(Q, P, Kqp)     (q x n doubles materialised)
//   :243-251           gradient rows N_i += alpha_j * k~(d_ij) * (q_i - p_j)   (un-normalised, :250)
//   :248, :302-303, :350-351   kernel map over Kqp
//   :252, :305, :353   F = Kqp * alpha
// and create<true>()'s normals at the training points (:166-181) when queries = training points.
// Outputs are zero-initialised (the reference accumulates into uninitialised memory).
//
// Two thread mappings, chosen by the host from the batch size:
//   * thread-per-query : training points broadcast from shared memory, n sequential terms per thread;
//   * warp-per-query   : lanes stride over the training points, warp-shuffle reduction — for the
//                        small batches (down to q = 1) that the reference's callers issue.
#include "gpr_common.cuh"
#include "gpr_kernels.h"

namespace gpr {

struct PredictArgs {
    const double* px; const double* py; const double* pz; const double* alpha;   // padded to N
    int n, N;
    const double* qx; const double* qy; const double* qz;
    int q;                 // real queries in this batch
    double* f;             // q
    double* grad;          // q x 3 column-major with leading dimension grad_ld, or null
    size_t grad_ld;
    double* panel;         // N x panel_ld: element (query, k) at panel[k*panel_ld + query], or null
    size_t panel_ld;       // >= q rounded up to 128; padded queries and padded k are written as 0
    double* part;          // split mode: per-chunk partial sums, [chunk][4][part_ld] (f, gx, gy, gz)
    size_t part_ld;
    int chunks_per_cta;    // training chunks (of PCHUNK points) handled by one CTA along blockIdx.y
    KernParams kp;
};

constexpr int PCHUNK = 1024;   // training points staged per shared-memory chunk

// Summation order (the same whatever the launch geometry, so that splitting the query set or the training
// set over CTAs never changes a bit): within a chunk of 1024 training points sequentially from zero, then
// the chunk sums sequentially in chunk order.  SPLIT: the chunk sums go to `part` and are added up by
// predict_reduce_kernel; otherwise the CTA walks all chunks and adds them up itself.
template <int KIND, bool GRAD, bool PANEL, bool SPLIT>
__global__ void __launch_bounds__(256) predict_thread_kernel(PredictArgs a) {
    __shared__ double4 sp[PCHUNK];   // x, y, z, alpha
    const int qi = blockIdx.x * 256 + threadIdx.x;
    const bool real = qi < a.q;
    const bool in_panel = PANEL && (size_t)qi < a.panel_ld;
    const double qx = real ? a.qx[qi] : 0.0, qy = real ? a.qy[qi] : 0.0, qz = real ? a.qz[qi] : 0.0;
    double f = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
    const int nchunks = (a.N + PCHUNK - 1) / PCHUNK;
    const int c0 = SPLIT ? blockIdx.y * a.chunks_per_cta : 0;
    const int c1 = SPLIT ? min(nchunks, c0 + a.chunks_per_cta) : nchunks;
    for (int ch = c0; ch < c1; ++ch) {
        const int base = ch * PCHUNK;
        __syncthreads();
        for (int k = threadIdx.x; k < PCHUNK; k += 256) {
            const int j = base + k;
            double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
            if (j < a.N) v = make_double4(a.px[j], a.py[j], a.pz[j], a.alpha[j]);
            sp[k] = v;
        }
        __syncthreads();
        const int lim = min(PCHUNK, a.N - base);
        double cf = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
#pragma unroll 4
        for (int k = 0; k < lim; ++k) {
            const double4 p = sp[k];
            const double dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
            const double d = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
            const double kv = kern_value<KIND>(a.kp, d);
            cf = fma(kv, p.w, cf);
            if (GRAD) {
                const double w = p.w * kern_diff<KIND>(a.kp, d, kv);
                cx = fma(w, dx, cx); cy = fma(w, dy, cy); cz = fma(w, dz, cz);
            }
            if (PANEL) {
                if (in_panel) a.panel[(size_t)(base + k) * a.panel_ld + qi] = (real && base + k < a.n) ? kv : 0.0;
            }
        }
        if (SPLIT) {
            if (real) {
                double* pp = a.part + (size_t)ch * 4 * a.part_ld + qi;
                pp[0] = cf;
                if (GRAD) { pp[a.part_ld] = cx; pp[2 * a.part_ld] = cy; pp[3 * a.part_ld] = cz; }
            }
        } else {
            f += cf;
            if (GRAD) { gx += cx; gy += cy; gz += cz; }
        }
    }
    if (!SPLIT && real) {
        a.f[qi] = f;
        if (GRAD) { a.grad[qi] = gx; a.grad[a.grad_ld + qi] = gy; a.grad[2 * a.grad_ld + qi] = gz; }
    }
}

template <bool GRAD>
__global__ void __launch_bounds__(256) predict_reduce_kernel(PredictArgs a, int nchunks) {
    const int qi = blockIdx.x * 256 + threadIdx.x;
    if (qi >= a.q) return;
    double f = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) {
        const double* pp = a.part + (size_t)ch * 4 * a.part_ld + qi;
        f += pp[0];
        if (GRAD) { gx += pp[a.part_ld]; gy += pp[2 * a.part_ld]; gz += pp[3 * a.part_ld]; }
    }
    a.f[qi] = f;
    if (GRAD) { a.grad[qi] = gx; a.grad[a.grad_ld + qi] = gy; a.grad[2 * a.grad_ld + qi] = gz; }
}

template <int KIND, bool GRAD, bool PANEL>
__global__ void __launch_bounds__(256) predict_warp_kernel(PredictArgs a) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (qi >= a.q) return;
    const double qx = a.qx[qi], qy = a.qy[qi], qz = a.qz[qi];
    double f = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
    const int jend = PANEL ? a.N : a.n;     // padded training points have alpha = 0 and coordinates 0
    for (int j = lane; j < jend; j += 32) {
        const double al = a.alpha[j];
        const double dx = qx - a.px[j], dy = qy - a.py[j], dz = qz - a.pz[j];
        const double d = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
        const double kv = kern_value<KIND>(a.kp, d);
        f = fma(kv, al, f);
        if (GRAD) {
            const double w = al * kern_diff<KIND>(a.kp, d, kv);
            gx = fma(w, dx, gx); gy = fma(w, dy, gy); gz = fma(w, dz, gz);
        }
        if (PANEL) a.panel[(size_t)j * a.panel_ld + qi] = j < a.n ? kv : 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        f += __shfl_xor_sync(0xffffffffu, f, o);
        if (GRAD) {
            gx += __shfl_xor_sync(0xffffffffu, gx, o);
            gy += __shfl_xor_sync(0xffffffffu, gy, o);
            gz += __shfl_xor_sync(0xffffffffu, gz, o);
        }
    }
    if (lane == 0) {
        a.f[qi] = f;
        if (GRAD) { a.grad[qi] = gx; a.grad[a.grad_ld + qi] = gy; a.grad[2 * a.grad_ld + qi] = gz; }
    }
}

template <int KIND, bool GRAD, bool PANEL>
static void launch_thread(const PredictArgs& a, int split, cudaStream_t st) {
    const size_t span = PANEL ? a.panel_ld : (size_t)a.q;
    const int qblocks = (int)((span + 255) / 256);
    if (split > 1) {
        dim3 grid(qblocks, split);
        predict_thread_kernel<KIND, GRAD, PANEL, true><<<grid, 256, 0, st>>>(a);
        predict_reduce_kernel<GRAD><<<(a.q + 255) / 256, 256, 0, st>>>(a, (a.N + PCHUNK - 1) / PCHUNK);
    } else {
        predict_thread_kernel<KIND, GRAD, PANEL, false><<<qblocks, 256, 0, st>>>(a);
    }
}

template <int KIND>
static cudaError_t launch_kind(const PredictArgs& a, bool warp_mode, int split, cudaStream_t st) {
    const bool grad = a.grad != nullptr, panel = a.panel != nullptr;
    if (warp_mode) {
        const int grid = (a.q + 7) / 8;
        if (grad && panel) predict_warp_kernel<KIND, true, true><<<grid, 256, 0, st>>>(a);
        else if (grad) predict_warp_kernel<KIND, true, false><<<grid, 256, 0, st>>>(a);
        else if (panel) predict_warp_kernel<KIND, false, true><<<grid, 256, 0, st>>>(a);
        else predict_warp_kernel<KIND, false, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (grad && panel) launch_thread<KIND, true, true>(a, split, st);
        else if (grad) launch_thread<KIND, true, false>(a, split, st);
        else if (panel) launch_thread<KIND, false, true>(a, split, st);
        else launch_thread<KIND, false, false>(a, split, st);
    }
    return cudaGetLastError();
}

// Number of CTAs along the training-point axis for the thread-per-query kernel: 1 when the queries alone
// fill the GPU, otherwise enough to reach ~4 CTAs per SM.
int predict_split(int q_span, int N, int num_sms) {
    const int qblocks = (q_span + 255) / 256;
    const int nchunks = (N + PCHUNK - 1) / PCHUNK;
    if (qblocks >= 2 * num_sms || nchunks < 2) return 1;
    int s = (4 * num_sms + qblocks - 1) / qblocks;
    return s > nchunks ? nchunks : s;
}
size_t predict_part_doubles(int q_span, int N) {
    return (size_t)((N + PCHUNK - 1) / PCHUNK) * 4 * (size_t)((q_span + 255) / 256 * 256);
}

cudaError_t launch_predict(const double* px, const double* py, const double* pz, const double* alpha, int n, int N,
                           const double* qx, const double* qy, const double* qz, int q, double* f, double* grad,
                           size_t grad_ld, double* panel, size_t panel_ld, const KernParams& kp, int warp_mode,
                           double* part, int split, cudaStream_t st) {
    if (q <= 0) return cudaSuccess;
    PredictArgs a;
    a.px = px; a.py = py; a.pz = pz; a.alpha = alpha; a.n = n; a.N = N;
    a.qx = qx; a.qy = qy; a.qz = qz; a.q = q; a.f = f; a.grad = grad; a.grad_ld = grad_ld;
    a.panel = panel; a.panel_ld = panel_ld; a.kp = kp;
    if (!part) split = 1;
    const int nchunks = (N + PCHUNK - 1) / PCHUNK;
    a.part = part; a.part_ld = (size_t)((q + 255) / 256 * 256);
    a.chunks_per_cta = split > 1 ? (nchunks + split - 1) / split : nchunks;
    if (split > 1) split = (nchunks + a.chunks_per_cta - 1) / a.chunks_per_cta;
    switch (kp.kind) {
        case 0: return launch_kind<0>(a, warp_mode != 0, split, st);
        case 1: return launch_kind<1>(a, warp_mode != 0, split, st);
        default: return launch_kind<2>(a, warp_mode != 0, split, st);
    }
}

// ---------------------------------------------------------------------------------------------
// computeTangentBasis for every row of the gradient (gp_regressor.hpp:29-44, :204-211).
// grad, N?, Tx, Ty are q x 3 column-major with leading dimension ld.
// ---------------------------------------------------------------------------------------------
__global__ void tangent_basis_kernel(const double* grad, size_t ld, int q, double* Tx, double* Ty) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    double g0 = grad[i], g1 = grad[ld + i], g2 = grad[2 * ld + i];
    double nrm = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
    double n0 = g0, n1 = g1, n2 = g2;
    if (nrm > 0.0) { n0 = g0 / nrm; n1 = g1 / nrm; n2 = g2 / nrm; }                  // :31
    // isApprox(UnitX, 1e-3): |N - X|^2 <= 1e-6 * min(|N|^2, 1)                        // :32
    double nn = n0 * n0 + n1 * n1 + n2 * n2;
    double d2 = (n0 - 1.0) * (n0 - 1.0) + n1 * n1 + n2 * n2;
    bool approx_x = d2 <= 1e-6 * fmin(nn, 1.0);
    double e0 = approx_x ? 0.0 : 1.0, e1 = approx_x ? 1.0 : 0.0;                     // :33 / :39
    double ne = n0 * e0 + n1 * e1;
    double t0 = e0 - n0 * ne, t1 = e1 - n1 * ne, t2 = -n2 * ne;
    double tn = sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    if (tn > 0.0) { t0 /= tn; t1 /= tn; t2 /= tn; }                                  // :34 / :40
    double u0 = n1 * t2 - n2 * t1, u1 = n2 * t0 - n0 * t2, u2 = n0 * t1 - n1 * t0;   // :35 / :41
    double un = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
    if (un > 0.0) { u0 /= un; u1 /= un; u2 /= un; }                                  // :36 / :42
    Tx[i] = t0; Tx[ld + i] = t1; Tx[2 * ld + i] = t2;
    Ty[i] = u0; Ty[ld + i] = u1; Ty[2 * ld + i] = u2;
}

cudaError_t launch_tangent_basis(const double* grad, size_t ld, int q, double* Tx, double* Ty, cudaStream_t st) {
    if (q <= 0) return cudaSuccess;
    tangent_basis_kernel<<<(q + 255) / 256, 256, 0, st>>>(grad, ld, q, Tx, Ty);
    return cudaGetLastError();
}

// Row-normalise a q x 3 column-major matrix in place (create<true>() normals, gp_regressor.hpp:174).
__global__ void normalize_rows_kernel(double* g, size_t ld, int q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    double a = g[i], b = g[ld + i], c = g[2 * ld + i];
    double n = sqrt(a * a + b * b + c * c);
    if (n > 0.0) { g[i] = a / n; g[ld + i] = b / n; g[2 * ld + i] = c / n; }
}

cudaError_t launch_normalize_rows(double* g, size_t ld, int q, cudaStream_t st) {
    if (q <= 0) return cudaSuccess;
    normalize_rows_kernel<<<(q + 255) / 256, 256, 0, st>>>(g, ld, q);
    return cudaGetLastError();
}

}  // namespace gpr

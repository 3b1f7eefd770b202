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
#include <climits>
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
    int n_panel;           // training points whose k* goes into the panel (the rest is written as 0)
    double* part;          // split mode: per-chunk partial sums, [chunk][4][part_ld] (f, gx, gy, gz)
    size_t part_ld;
    int chunks_per_cta;    // training chunks (of PCHUNK points) handled by one CTA along blockIdx.y
    KernParams kp;
};

constexpr int PCHUNK = 1024;   // training points staged per shared-memory chunk
constexpr int SMALL_MS = 8;    // at most this many mean blocks in the small-batch kernel

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
                if (in_panel) a.panel[(size_t)(base + k) * a.panel_ld + qi] = (real && base + k < a.n_panel) ? kv : 0.0;
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
        if (PANEL) a.panel[(size_t)j * a.panel_ld + qi] = j < a.n_panel ? kv : 0.0;
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
                           size_t grad_ld, double* panel, size_t panel_ld, int n_panel, const KernParams& kp,
                           int warp_mode, double* part, int split, cudaStream_t st) {
    if (q <= 0) return cudaSuccess;
    PredictArgs a;
    a.px = px; a.py = py; a.pz = pz; a.alpha = alpha; a.n = n; a.N = N;
    a.qx = qx; a.qy = qy; a.qz = qz; a.q = q; a.f = f; a.grad = grad; a.grad_ld = grad_ld;
    a.panel = panel; a.panel_ld = panel_ld; a.n_panel = n_panel < n ? n_panel : n; a.kp = kp;
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
__device__ __forceinline__ void tangent_basis_dev(double g0, double g1, double g2, double (&T)[3], double (&U)[3]) {
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
    T[0] = t0; T[1] = t1; T[2] = t2; U[0] = u0; U[1] = u1; U[2] = u2;
}

__global__ void tangent_basis_kernel(const double* grad, size_t ld, int q, double* Tx, double* Ty) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    double T[3], U[3];
    tangent_basis_dev(grad[i], grad[ld + i], grad[2 * ld + i], T, U);
    Tx[i] = T[0]; Tx[ld + i] = T[1]; Tx[2 * ld + i] = T[2];
    Ty[i] = U[0]; Ty[ld + i] = U[1]; Ty[2 * ld + i] = U[2];
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

// ---------------------------------------------------------------------------------------------
// Fused small-batch predict (q <= 8): the call pattern of every real caller of the reference — ONE query
// per evaluate() from hundreds of host threads (src/gp_node.cpp:1027-1038, :1074; include/atlas/
// atlas_variance.hpp:78, :201; include/atlas/atlas.hpp:225, :259).  One launch, no device-side staging:
//   * queries are read from, and all results written to, a pinned host buffer mapped into the device
//     address space (hio): no cudaMemcpy in the call;
//   * variance blocks: v = X k* as a matrix-vector product over the triangle of X = L^-1 (bandwidth bound,
//     4 n^2 bytes); k* is evaluated on the fly per 256-point chunk into shared memory (never stored);
//   * mean blocks: f and the un-normalised gradient over a slice of the training points;
//   * the last block to finish (atomic ticket) reduces the partial results in a fixed order, applies
//     computeTangentBasis if asked, writes the outputs and re-arms the ticket.
// hio layout (doubles): [0,8) qx [8,16) qy [16,24) qz | [24,32) f [32,40) var [40,64) grad (c*8+i)
//                       [64,88) tx [88,112) ty
// ---------------------------------------------------------------------------------------------
struct SmallArgs {
    const double* px; const double* py; const double* pz; const double* alpha;
    int n, N;
    const double* X; size_t ld;     // L^-1 (null when no variance is wanted)
    double* hio;                    // mapped pinned host buffer, SMALL_HIO_DOUBLES doubles
    unsigned int* ticket;           // global ticket
    unsigned int* rowticket;        // one per 64-row block
    double* fpart;                  // SMALL_MS * 8 * 4 (mean / gradient partial sums)
    double* rowpart;                // (N/64) * 8: squared norm of each 64-row block of v
    double* part;                   // vs * N * 8 (partial dot products per k-split)
    int q, want_var, want_grad, want_t;
    int vs, kspan;                  // k-splits and their extent (multiple of 256)
    int nvar_blocks, nmean_blocks;
    // indefinite tail block (gpr_tail.cu): X covers the leading n_var points / Nv padded rows; the last n_tail points
    // enter through var -= w^T S^-1 w, w = k_2 - Z^T k_1.  n_tail == 0: n_var == n, Nv == N.
    int n_var, Nv, n_tail, mp;
    const double* tZ; const double* tSinv;
    double* wpart;                  // SMALL_MS * 8 * 256: partial Z^T k_1 per mean block and query
    double k0;
    KernParams kp;
};

template <int KIND, int Q>
__global__ void __launch_bounds__(256) predict_small_kernel(SmallArgs a) {
    __shared__ double ks[256][Q];
    __shared__ double red[256 * (Q > 1 ? Q : 1)];
    __shared__ double sq[3][8];
    __shared__ unsigned int s_last;
    const int tid = threadIdx.x;
    if (tid < 24) sq[tid >> 3][tid & 7] = (tid & 7) < a.q ? a.hio[tid] : 0.0;
    __syncthreads();
    bool to_global = false;          // does this block take a global ticket?
    if ((int)blockIdx.x < a.nvar_blocks) {
        // ---- variance block: 64 rows x one k-split; thread = (row rl, k-phase kq) ----
        const int bx = blockIdx.x / a.vs, by = blockIdx.x % a.vs;
        const int rl = tid & 63, kq = tid >> 6;
        const int r = bx * 64 + rl;
        const int kbeg = by * a.kspan;
        const int kend = min(min(kbeg + a.kspan, a.Nv), bx * 64 + 64);
        if (kbeg >= kend) return;                                    // above the diagonal: nothing to do
        const int nsplit = (min(a.Nv, bx * 64 + 64) + a.kspan - 1) / a.kspan;   // non-empty splits of this row block
        double acc[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) acc[i] = 0.0;
        for (int c0 = kbeg; c0 < kend; c0 += 256) {
            __syncthreads();
            {
                const int c = c0 + tid;
                const bool real = c < a.n_var;
                const double x = real ? a.px[c] : 0.0, y = real ? a.py[c] : 0.0, z = real ? a.pz[c] : 0.0;
#pragma unroll
                for (int i = 0; i < Q; ++i) {
                    const double dx = sq[0][i] - x, dy = sq[1][i] - y, dz = sq[2][i] - z;
                    const double d = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
                    ks[tid][i] = real ? kern_value<KIND>(a.kp, d) : 0.0;
                }
            }
            __syncthreads();
            // this thread's 64 k values of the chunk, 16 loads in flight at a time
            const int kl0 = 64 * kq;
            const double* xr = a.X + (size_t)(c0 + kl0) * a.ld + r;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double xv[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int c = c0 + kl0 + 16 * b + u;
                    xv[u] = (c <= r && c < kend) ? __ldcs(xr + (size_t)(16 * b + u) * a.ld) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
#pragma unroll
                    for (int i = 0; i < Q; ++i) acc[i] = fma(xv[u], ks[kl0 + 16 * b + u][i], acc[i]);
                }
            }
        }
        // sum the four k-phases in a fixed order
        __syncthreads();
        double* r4 = &ks[0][0];                      // reuse: [kq][rl][i], 4*64*Q doubles = 256*Q
#pragma unroll
        for (int i = 0; i < Q; ++i) r4[(kq * 64 + rl) * Q + i] = acc[i];
        __syncthreads();
        if (kq == 0) {
#pragma unroll
            for (int i = 0; i < Q; ++i)
                a.part[((size_t)by * a.Nv + r) * 8 + i] = ((r4[rl * Q + i] + r4[(64 + rl) * Q + i]) + r4[(128 + rl) * Q + i]) + r4[(192 + rl) * Q + i];
        }
        // row-block ticket: the last of the nsplit blocks squares and sums the 64 rows
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(a.rowticket + bx, 1u) == (unsigned)(nsplit - 1) ? 1u : 0u;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (tid < 64) {
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                double v = 0.0;
                for (int y = 0; y < nsplit; ++y) v += __ldcg(a.part + ((size_t)y * a.Nv + bx * 64 + tid) * 8 + i);
                red[tid * Q + i] = v * v;
            }
        }
        __syncthreads();
        if (tid < Q) {
            double ssum = 0.0;
            for (int rr = 0; rr < 64; ++rr) ssum += red[rr * Q + tid];
            a.rowpart[(size_t)bx * 8 + tid] = ssum;
        }
        if (tid == 0) a.rowticket[bx] = 0u;
        to_global = true;
    } else {
        // ---- mean block: f and gradient over a slice of the training points ----
        const int mb = blockIdx.x - a.nvar_blocks;
        const int span = (a.n + a.nmean_blocks - 1) / a.nmean_blocks;
        const int j0 = mb * span, j1 = min(a.n, j0 + span);
        for (int i = 0; i < a.q; ++i) {
            double f = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
            for (int j = j0 + tid; j < j1; j += 256) {
                const double al = a.alpha[j];
                const double dx = sq[0][i] - a.px[j], dy = sq[1][i] - a.py[j], dz = sq[2][i] - a.pz[j];
                const double d = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
                const double kv = kern_value<KIND>(a.kp, d);
                f = fma(kv, al, f);
                const double w = al * kern_diff<KIND>(a.kp, d, kv);
                gx = fma(w, dx, gx); gy = fma(w, dy, gy); gz = fma(w, dz, gz);
            }
            if (a.n_tail > 0 && a.want_var) {
                // partial W = Z^T k_1 over this block's slice of the leading points: chunks of 256 kernel values in
                // shared memory, thread = (tail column ta, row phase), phases summed in a fixed order
                const int nph = 256 / a.mp, ta = tid % a.mp, ph = tid / a.mp;
                const int jv1 = min(j1, a.n_var);
                double wacc = 0.0;
                for (int c0 = j0; c0 < jv1; c0 += 256) {
                    __syncthreads();
                    {
                        const int j = c0 + tid;
                        double kv = 0.0;
                        if (j < jv1) {
                            const double dx = sq[0][i] - a.px[j], dy = sq[1][i] - a.py[j], dz = sq[2][i] - a.pz[j];
                            kv = kern_value<KIND>(a.kp, sqrt(fma(dz, dz, fma(dy, dy, dx * dx))));
                        }
                        ks[tid][0] = kv;
                    }
                    __syncthreads();
                    if (ph < nph) {
                        const int lim = min(256, jv1 - c0);
                        const double* zc = a.tZ + ((size_t)(ta >> 5) * a.ld + c0) * 32 + (ta & 31);
                        for (int jl = ph; jl < lim; jl += nph) wacc = fma(zc[(size_t)jl * 32], ks[jl][0], wacc);
                    }
                }
                __syncthreads();
                red[tid] = ph < nph ? wacc : 0.0;
                __syncthreads();
                if (tid < a.mp) {
                    double t = 0.0;
                    for (int p2 = 0; p2 < nph; ++p2) t += red[p2 * a.mp + tid];
                    a.wpart[((size_t)mb * 8 + i) * 256 + tid] = t;
                }
                __syncthreads();
            }
            double vals[4] = {f, gx, gy, gz};
            for (int c = 0; c < (a.want_grad ? 4 : 1); ++c) {
                // fixed-order tree: warp shuffles, then the 8 warp sums
                double v = vals[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                __syncthreads();
                if ((tid & 31) == 0) red[tid >> 5] = v;
                __syncthreads();
                if (tid == 0) {
                    double t = 0.0;
                    for (int w8 = 0; w8 < 8; ++w8) t += red[w8];
                    a.fpart[(mb * 8 + i) * 4 + c] = t;
                }
            }
        }
        to_global = true;
    }
    if (!to_global) return;
    // ---- global ticket: one per non-empty row block + one per mean block; the last one finishes ----
    __threadfence();
    __syncthreads();
    const unsigned total = (unsigned)(a.want_var ? (a.Nv + 63) / 64 : 0) + (unsigned)a.nmean_blocks;
    if (tid == 0) s_last = atomicAdd(a.ticket, 1u) == total - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double* out = a.hio + 24;
    const int nrb = (a.Nv + 63) / 64;
    for (int i = 0; i < a.q; ++i) {
        if (a.want_var) {
            // squared norm = sum over the row blocks, fixed order: strided partial sums, then a tree
            double sacc = 0.0;
            for (int b = tid; b < nrb; b += 256) sacc += __ldcg(a.rowpart + (size_t)b * 8 + i);
            __syncthreads();
            red[tid] = sacc;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if (tid < o) red[tid] += red[tid + o];
                __syncthreads();
            }
            double var = a.k0 - red[0];
            if (a.n_tail > 0) {
                // var -= w^T S^-1 w,  w = k_2 - Z^T k_1
                double* wsm = &ks[0][0];
                __syncthreads();
                if (tid < a.mp) {
                    double wv = 0.0;
                    if (tid < a.n_tail) {
                        double W = 0.0;
                        for (int mb = 0; mb < a.nmean_blocks; ++mb) W += __ldcg(a.wpart + ((size_t)mb * 8 + i) * 256 + tid);
                        const int j = a.n_var + tid;
                        const double dx = sq[0][i] - a.px[j], dy = sq[1][i] - a.py[j], dz = sq[2][i] - a.pz[j];
                        wv = kern_value<KIND>(a.kp, sqrt(fma(dz, dz, fma(dy, dy, dx * dx)))) - W;
                    }
                    wsm[tid] = wv;
                }
                __syncthreads();
                double term = 0.0;
                if (tid < a.n_tail) {
                    double row = 0.0;
                    for (int b = 0; b < a.n_tail; ++b) row = fma(a.tSinv[(size_t)tid * a.mp + b], wsm[b], row);
                    term = wsm[tid] * row;
                }
                __syncthreads();
                red[tid] = term;
                __syncthreads();
                for (int o = 128; o > 0; o >>= 1) {
                    if (tid < o) red[tid] += red[tid + o];
                    __syncthreads();
                }
                var -= red[0];
            }
            if (tid == 0) out[8 + i] = var;
        }
        if (tid == 0) {
            double v[4] = {0.0, 0.0, 0.0, 0.0};
            for (int c = 0; c < (a.want_grad ? 4 : 1); ++c)
                for (int mb = 0; mb < a.nmean_blocks; ++mb) v[c] += __ldcg(a.fpart + (mb * 8 + i) * 4 + c);
            out[i] = v[0];
            if (a.want_grad) {
                out[16 + i] = v[1]; out[16 + 8 + i] = v[2]; out[16 + 16 + i] = v[3];
                if (a.want_t) {
                    double T[3], U[3];
                    tangent_basis_dev(v[1], v[2], v[3], T, U);
                    for (int c = 0; c < 3; ++c) { out[40 + c * 8 + i] = T[c]; out[64 + c * 8 + i] = U[c]; }
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) { __threadfence_system(); *a.ticket = 0u; }
}

template <int KIND>
static void launch_small_q(const SmallArgs& a, int grid, cudaStream_t st) {
    if (a.q <= 1) predict_small_kernel<KIND, 1><<<grid, 256, 0, st>>>(a);
    else if (a.q <= 2) predict_small_kernel<KIND, 2><<<grid, 256, 0, st>>>(a);
    else if (a.q <= 4) predict_small_kernel<KIND, 4><<<grid, 256, 0, st>>>(a);
    else predict_small_kernel<KIND, 8><<<grid, 256, 0, st>>>(a);
}

static int small_vs(int N) { int v = (N + 1023) / 1024; return v < 1 ? 1 : (v > 16 ? 16 : v); }

constexpr int SMALL_WPART = SMALL_MS * 8 * 256;

// ticket (2) | fpart | wpart | row tickets (one unsigned per 64 rows) | rowpart | part
size_t predict_small_scratch_doubles(int Nv) {
    const size_t nrb = (size_t)(Nv + 63) / 64;
    return 2 + SMALL_MS * 8 * 4 + SMALL_WPART + (nrb + 1) / 2 + nrb * 8 + (size_t)small_vs(Nv) * Nv * 8;
}

// scratch: predict_small_scratch_doubles(Nv) doubles, zero-initialised when allocated AND whenever Nv changes (the
// tickets inside it are re-armed by the kernel itself).  hio: device pointer of the mapped host buffer.
// n / N: all training points (mean, gradient); n_var / Nv: the points / padded rows covered by X = L^-1 (== n / N unless
// the model has an indefinite tail block of n_tail = n - n_var points, described by tZ / tSinv / mp).
cudaError_t launch_predict_small(const double* px, const double* py, const double* pz, const double* alpha, int n, int N,
                                 const double* X, size_t ld, double* hio, double* scratch, int q, int want_var,
                                 int want_grad, int want_t, double k0, const KernParams& kp, int n_var, int Nv, int mp,
                                 const double* tZ, const double* tSinv, cudaStream_t st) {
    if (q <= 0 || q > 8) return cudaErrorInvalidValue;
    SmallArgs a;
    a.px = px; a.py = py; a.pz = pz; a.alpha = alpha; a.n = n; a.N = N; a.X = X; a.ld = ld; a.hio = hio;
    a.n_var = n_var; a.Nv = Nv; a.n_tail = n - n_var; a.mp = mp > 0 ? mp : 32; a.tZ = tZ; a.tSinv = tSinv;
    const size_t nrb = (size_t)(Nv + 63) / 64;
    a.ticket = reinterpret_cast<unsigned int*>(scratch);
    a.fpart = scratch + 2;
    a.wpart = a.fpart + SMALL_MS * 8 * 4;
    a.rowticket = reinterpret_cast<unsigned int*>(a.wpart + SMALL_WPART);
    a.rowpart = a.wpart + SMALL_WPART + (nrb + 1) / 2;
    a.part = a.rowpart + nrb * 8;
    a.q = q; a.want_var = want_var && X != nullptr; a.want_grad = want_grad; a.want_t = want_t; a.k0 = k0; a.kp = kp;
    a.vs = small_vs(Nv);
    a.kspan = ((Nv + a.vs - 1) / a.vs + 255) / 256 * 256;
    a.nvar_blocks = a.want_var ? (int)nrb * a.vs : 0;
    a.nmean_blocks = (n + 2047) / 2048 < SMALL_MS ? (n + 2047) / 2048 : SMALL_MS;
    if (a.nmean_blocks < 1) a.nmean_blocks = 1;
    const int grid = a.nvar_blocks + a.nmean_blocks;
    switch (kp.kind) {
        case 0: launch_small_q<0>(a, grid, st); break;
        case 1: launch_small_q<1>(a, grid, st); break;
        default: launch_small_q<2>(a, grid, st); break;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Batched iso-surface sampling: the node's fakeDeterministicSampling / samplePoint
// (src/gp_node.cpp:998-1100) walks a lattice with one thread and one evaluate(q = 1) per lattice point and
// keeps the points with |f| <= 0.01.  Here the lattice is generated on the device chunk by chunk, the fused
// mean kernel evaluates it, and the survivors are compacted on the device; the variance (n^2 flop per
// point) is then computed for the survivors only.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_fill_kernel(const double* __restrict__ axis, int na, unsigned long long g0,
                                                        int count, double* qx, double* qy, double* qz) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    const unsigned long long g = g0 + i;              // x-major, then y, then z (the node's loop nest)
    const int iz = (int)(g % na), iy = (int)((g / na) % na), ix = (int)(g / ((unsigned long long)na * na));
    qx[i] = axis[ix]; qy[i] = axis[iy]; qz[i] = axis[iz];
}

__global__ void __launch_bounds__(256) grid_select_kernel(const double* __restrict__ f, unsigned long long g0, int count,
                                                          double tol, unsigned int* counter, unsigned long long* sel_idx,
                                                          double* sel_f) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    const double v = f[i];
    if (fabs(v) <= tol) {                              // src/gp_node.cpp:1075
        const unsigned int pos = atomicAdd(counter, 1u);
        sel_idx[pos] = g0 + i;
        sel_f[pos] = v;
    }
}

cudaError_t launch_grid_fill(const double* axis, int na, unsigned long long g0, int count, double* qx, double* qy,
                             double* qz, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    grid_fill_kernel<<<(count + 255) / 256, 256, 0, st>>>(axis, na, g0, count, qx, qy, qz);
    return cudaGetLastError();
}

cudaError_t launch_grid_select(const double* f, unsigned long long g0, int count, double tol, unsigned int* counter,
                               unsigned long long* sel_idx, double* sel_f, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    grid_select_kernel<<<(count + 255) / 256, 256, 0, st>>>(f, g0, count, tol, counter, sel_idx, sel_f);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Batched AtlasVariance::sampleOnChart (include/atlas/atlas_variance.hpp:147-219): uniform annulus samples on the
// tangent disc of each chart, pK = Tkl * (R sqrt(r) cos th, R sqrt(r) sin th, 0, 1) with Tkl = [Tx Ty N C] (:166-195),
// then — after the batched mean + variance of all samples — the per-chart order by decreasing variance (:214-218).
// frames: 13 doubles per chart (C, N, Tx, Ty, R); offsets: n_charts + 1 prefix sums of the sample counts.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chart_fill_kernel(const double* __restrict__ frames, const unsigned long long* __restrict__ offsets,
                                                         int n_charts, const double* __restrict__ r, const double* __restrict__ th,
                                                         int total, double* qx, double* qy, double* qz) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    int lo = 0, hi = n_charts;                         // chart c with offsets[c] <= i < offsets[c+1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offsets[mid] <= (unsigned long long)i) lo = mid; else hi = mid; }
    const double* F = frames + 13 * (size_t)lo;
    const double R = F[12];
    const double sr = sqrt(r[i]);
    const double a = __dmul_rn(__dmul_rn(R, sr), cos(th[i]));
    const double b = __dmul_rn(__dmul_rn(R, sr), sin(th[i]));
    double out[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)                        // row c of Tkl times (a, b, 0, 1), accumulated left to right
        out[c] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(F[6 + c], a), __dmul_rn(F[9 + c], b)), __dmul_rn(F[3 + c], 0.0)), F[c]);
    qx[i] = out[0]; qy[i] = out[1]; qz[i] = out[2];
}

// One CTA per chart: rank of every sample by (variance descending, index ascending); flags non-finite f / v.
__global__ void __launch_bounds__(256) chart_rank_kernel(const double* __restrict__ f, const double* __restrict__ v,
                                                         const unsigned long long* __restrict__ offsets,
                                                         unsigned long long* order, int* bad) {
    const unsigned long long o0 = offsets[blockIdx.x], o1 = offsets[blockIdx.x + 1];
    const int cnt = (int)(o1 - o0);
    for (int i = threadIdx.x; i < cnt; i += 256) {
        const double vi = v[o0 + i];
        if (!isfinite(vi) || !isfinite(f[o0 + i])) atomicExch(bad, 1);
        int rank = 0;
        for (int j = 0; j < cnt; ++j) {
            const double vj = v[o0 + j];
            rank += (vj > vi) || (vj == vi && j < i);
        }
        order[o0 + rank] = (unsigned long long)i;
    }
}

cudaError_t launch_chart_fill(const double* frames, const unsigned long long* offsets, int n_charts, const double* r,
                              const double* th, int total, double* qx, double* qy, double* qz, cudaStream_t st) {
    if (total <= 0) return cudaSuccess;
    chart_fill_kernel<<<(total + 255) / 256, 256, 0, st>>>(frames, offsets, n_charts, r, th, total, qx, qy, qz);
    return cudaGetLastError();
}

cudaError_t launch_chart_rank(const double* f, const double* v, const unsigned long long* offsets, int n_charts,
                              unsigned long long* order, int* bad, cudaStream_t st) {
    if (n_charts <= 0) return cudaSuccess;
    chart_rank_kernel<<<n_charts, 256, 0, st>>>(f, v, offsets, order, bad);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Batched projection onto the iso-surface f = 0: AtlasBase::project
// (include/atlas/atlas.hpp:201-276) is a fixed-step gradient descent that calls evaluate(q = 1) twice per
// iteration, up to 500 iterations, for ONE point.  Here one CTA owns one point and runs the whole iteration
// on the device (f and the un-normalised gradient over all training points per iteration, block-reduced in
// a fixed order); any number of points are projected by one launch.  Same update rule and the same three
// stopping criteria; the variance the reference also computes in every iteration is only printed there and
// is not computed here.
// status: > 0 iterations used when a criterion was met (f_tol or improve_tol), -(max_iter) when the iteration
// budget ran out, INT_MIN if f became NaN/Inf (the reference throws "f is nan or inf").
// ---------------------------------------------------------------------------------------------
struct ProjectArgs {
    const double* px; const double* py; const double* pz; const double* alpha; int n;
    const double* x; const double* y; const double* z;       // start points
    const double* nx; const double* ny; const double* nz;    // initial (un-normalised) gradients
    int count;
    double f_tol, improve_tol, step_mul; int max_iter;
    double* ox; double* oy; double* oz; int* status;
    KernParams kp;
};

template <int KIND>
__device__ __forceinline__ void block_eval(const ProjectArgs& a, double qx, double qy, double qz, double (&out)[4],
                                           double (*red)[4]) {
    double f = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
    for (int j = threadIdx.x; j < a.n; j += 256) {
        const double al = a.alpha[j];
        const double dx = qx - a.px[j], dy = qy - a.py[j], dz = qz - a.pz[j];
        const double d = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
        const double kv = kern_value<KIND>(a.kp, d);
        f = fma(kv, al, f);
        const double w = al * kern_diff<KIND>(a.kp, d, kv);
        gx = fma(w, dx, gx); gy = fma(w, dy, gy); gz = fma(w, dz, gz);
    }
    double v[4] = {f, gx, gy, gz};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 4; ++c) red[threadIdx.x >> 5][c] = v[c];
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double t = 0.0;
        for (int w8 = 0; w8 < 8; ++w8) t += red[w8][c];
        out[c] = t;
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) project_kernel(ProjectArgs a) {
    __shared__ double red[8][4];
    const int b = blockIdx.x;
    if (b >= a.count) return;
    double cx = a.x[b], cy = a.y[b], cz = a.z[b];
    double gx = a.nx[b], gy = a.ny[b], gz = a.nz[b];
    double e[4];
    block_eval<KIND>(a, cx, cy, cz, e, red);                     // atlas.hpp:225
    double f_cur = e[0];
    int status = -a.max_iter;
    for (int iter = 0; iter < a.max_iter; ++iter) {
        if (isnan(f_cur) || isinf(f_cur)) { status = INT_MIN; break; }                 // :227-231
        if (fabs(f_cur) < a.f_tol) { status = iter + 1; break; }                        // :236-241
        const double sx = a.step_mul * f_cur * gx, sy = a.step_mul * f_cur * gy, sz = a.step_mul * f_cur * gz;   // :245
        const double ns2 = sx * sx + sy * sy + sz * sz;
        if (ns2 <= 1e4 && ns2 > 1e-12) { cx -= sx; cy -= sy; cz -= sz; }                // :246-251 (isMuchSmallerThan / isZero)
        block_eval<KIND>(a, cx, cy, cz, e, red);                                        // :259
        const double ng2 = e[1] * e[1] + e[2] * e[2] + e[3] * e[3];
        if (ng2 <= 1e4 && ng2 > 1e-10) { gx = e[1]; gy = e[2]; gz = e[3]; }            // :260-265
        if (fabs(e[0] - f_cur) < a.improve_tol) { status = iter + 1; break; }           // :266-271
        f_cur = e[0];
    }
    if (threadIdx.x == 0) { a.ox[b] = cx; a.oy[b] = cy; a.oz[b] = cz; a.status[b] = status; }
}

cudaError_t launch_project(const double* px, const double* py, const double* pz, const double* alpha, int n,
                           const double* xyz_in /*x|y|z|nx|ny|nz, each `ld` apart*/, size_t ld, int count, double f_tol,
                           double improve_tol, int max_iter, double step_mul, double* out /*x|y|z, ld apart*/,
                           int* status, const KernParams& kp, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    ProjectArgs a;
    a.px = px; a.py = py; a.pz = pz; a.alpha = alpha; a.n = n;
    a.x = xyz_in; a.y = xyz_in + ld; a.z = xyz_in + 2 * ld; a.nx = xyz_in + 3 * ld; a.ny = xyz_in + 4 * ld; a.nz = xyz_in + 5 * ld;
    a.count = count; a.f_tol = f_tol; a.improve_tol = improve_tol; a.step_mul = step_mul; a.max_iter = max_iter;
    a.ox = out; a.oy = out + ld; a.oz = out + 2 * ld; a.status = status; a.kp = kp;
    switch (kp.kind) {
        case 0: project_kernel<0><<<count, 256, 0, st>>>(a); break;
        case 1: project_kernel<1><<<count, 256, 0, st>>>(a); break;
        default: project_kernel<2><<<count, 256, 0, st>>>(a); break;
    }
    return cudaGetLastError();
}

}  // namespace gpr

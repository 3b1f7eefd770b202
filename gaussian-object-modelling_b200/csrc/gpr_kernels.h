// gpr_kernels.h — host-side launchers of the sm_100a kernels (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "gpr_common.cuh"

namespace gpr {

// K1 (gpr_cov.cu)
cudaError_t launch_cov_build(const double* x, const double* y, const double* z, const double* sigma2, int n, int nb,
                             int tile_row0, double* K, size_t ld, unsigned long long* rmax_bits, const KernParams& kp,
                             cudaStream_t st);
// K2 (gpr_factor.cu).  scratch: at least 4 + nb*nb ints.  peers: replicas (same ld) that receive every finished tile of
// L and Dinv through peer stores while the factorisation runs.
constexpr int MAX_CHOL_PEERS = 7;
struct CholPeers { int n; double* L[MAX_CHOL_PEERS]; double* Dinv[MAX_CHOL_PEERS]; };
cudaError_t launch_cholesky(double* A, size_t ld, int nb, double* Dinv, int* scratch, int num_sms, int serial,
                            cudaStream_t st, long long* trace = nullptr, const CholPeers* peers = nullptr);
// The same factorisation with the bulk of the flops on the INT8 tensor cores: panels of `panel_tiles` tile columns are factorised
// by the FP64 tile kernel, everything to the left of a panel is applied to it beforehand as one sliced-integer product
// (gpr_ozaki.cu MODE 1).  Ls: S * N * N bytes of slice workspace (N = nb * 128) followed by N doubles (the per-row scales)
// and the (N / 128) x (N / 64) byte map of all-zero slice blocks:
// ozaki_fit_workspace_bytes(S, N); ctrl: 2 ints.  The final last_tiles tile columns are factorised as one panel.
size_t ozaki_fit_workspace_bytes(int S, size_t N);
cudaError_t launch_cholesky_int8(double* A, size_t ld, int nb, double* Dinv, int* scratch, int num_sms, cudaStream_t st,
                                 const CholPeers* peers, signed char* Ls, int S, int panel_tiles, int last_tiles, int* ctrl);
cudaError_t launch_dinv_from_l(const double* L, size_t ld, int nb, double* Dinv, cudaStream_t st);
cudaError_t launch_linv(const double* L, double* X, size_t ld, int nb, const double* Dinv, int* scratch, int num_sms,
                        cudaStream_t st);
// K3 (gpr_solve.cu).  scratch: at least 4 + nb ints.
cudaError_t launch_trsv(int backward, const double* L, size_t ld, int nb, const double* Dinv, const double* rhs,
                        double* out, int* scratch, int num_sms, cudaStream_t st);
// Iterative refinement of alpha (gpr_solve.cu): r = y - K alpha in double-double; part: residual_scratch_doubles(N).
size_t residual_scratch_doubles(int N);
cudaError_t launch_residual(const double* xyz, size_t ld, const double* sigma2, const double* label, const double* alpha,
                            int n, int N, double* part, double* r, const KernParams& kp, cudaStream_t st);
cudaError_t launch_axpy1(double* a, const double* d, int n, cudaStream_t st);
cudaError_t launch_commit_if_clear(const int* flag, const double* src, double* dst, int n, cudaStream_t st);
// out = X^T (X in): K^-1 applied through the resident inverse factor (two bandwidth-bound triangular matrix-vector
// products, no dependency chain); scratch: (tri_gemv_splits(n) + 1) * n doubles.
int tri_gemv_splits(int n);
cudaError_t launch_solve_with_inverse(const double* X, size_t ld, int n, const double* in, double* out, double* scratch,
                                      cudaStream_t st);
// K4 (gpr_predict.cu).  part/split: optional split of the training points over CTAs for the thread-per-query
// kernel (split from predict_split, part of predict_part_doubles(q, N) doubles); bit-identical results either way.
int predict_split(int q_span, int N, int num_sms);
size_t predict_part_doubles(int q_span, int N);
cudaError_t launch_predict(const double* px, const double* py, const double* pz, const double* alpha, int n, int N,
                           const double* qx, const double* qy, const double* qz, int q, double* f, double* grad,
                           size_t grad_ld, double* panel, size_t panel_ld, int n_panel, const KernParams& kp,
                           int warp_mode, double* part, int split, cudaStream_t st);
// Fused q <= 8 path: one launch, queries and results in a mapped pinned host buffer (112 doubles, layout in
// gpr_predict.cu).  scratch: predict_small_scratch_doubles(N) doubles, zeroed once at allocation.
constexpr int SMALL_HIO_DOUBLES = 112;
size_t predict_small_scratch_doubles(int Nv);
cudaError_t launch_predict_small(const double* px, const double* py, const double* pz, const double* alpha, int n, int N,
                                 const double* X, size_t ld, double* hio, double* scratch, int q, int want_var,
                                 int want_grad, int want_t, double k0, const KernParams& kp, int n_var, int Nv, int mp,
                                 const double* tZ, const double* tSinv, cudaStream_t st);
// Lattice generation + |f| <= tol compaction for the batched iso-surface sampler.
cudaError_t launch_grid_fill(const double* axis, int na, unsigned long long g0, int count, double* qx, double* qy,
                             double* qz, cudaStream_t st);
cudaError_t launch_grid_select(const double* f, unsigned long long g0, int count, double tol, unsigned int* counter,
                               unsigned long long* sel_idx, double* sel_f, cudaStream_t st);
// Batched AtlasVariance::sampleOnChart: annulus samples on the charts' tangent discs; per-chart order by variance.
cudaError_t launch_chart_fill(const double* frames, const unsigned long long* offsets, int n_charts, const double* r,
                              const double* th, int total, double* qx, double* qy, double* qz, cudaStream_t st);
cudaError_t launch_chart_rank(const double* f, const double* v, const unsigned long long* offsets, int n_charts,
                              unsigned long long* order, int* bad, cudaStream_t st);
// Batched gradient-descent projection onto f = 0 (AtlasBase::project), one CTA per point.
cudaError_t launch_project(const double* px, const double* py, const double* pz, const double* alpha, int n,
                           const double* xyz_in, size_t ld, int count, double f_tol, double improve_tol, int max_iter,
                           double step_mul, double* out, int* status, const KernParams& kp, cudaStream_t st);
cudaError_t launch_tangent_basis(const double* grad, size_t ld, int q, double* Tx, double* Ty, cudaStream_t st);
cudaError_t launch_normalize_rows(double* g, size_t ld, int q, cudaStream_t st);
// K3' (gpr_var.cu)
cudaError_t launch_variance(const double* X, size_t ld, int nb, const double* panel, size_t panel_ld, int q,
                            double* partial, double k0, double* var, cudaStream_t st, size_t panel_pitch = 0);
// Same result without L^-1: blocked forward substitution V = L^-1 K*^T in place in the panel (panel is overwritten).
cudaError_t launch_variance_trsm(const double* L, size_t ld, int nb, const double* Dinv, double* panel, size_t panel_ld,
                                 int q, double* partial, double k0, double* var, cudaStream_t st);
cudaError_t launch_variance_small(const double* X, size_t ld, int N, const double* panel, size_t panel_ld, int q,
                                  double* part, double k0, double* var, cudaStream_t st);
cudaError_t launch_var_finalize(const double* partial, size_t panel_ld, int nb, int q, double k0, double* var, cudaStream_t st);
// K3'' (gpr_ozaki.cu): the variance product on the INT8 tensor cores (tcgen05 kind::i8 + TMEM + TMA), FP64-equivalent by slicing.
int ozaki_tile_n(int S);
long long ozaki_max_k(int S, int base254);
bool ozaki_supported(int S, int base254, long long k_extent);
cudaError_t launch_ozaki_product(const signed char* As, size_t a_pitch, size_t a_slice, int nrt, const signed char* Bs, size_t b_pitch,
                                 size_t b_slice, size_t b_rows, int q, size_t q_pad, size_t k_extent, int tri, int S, int base254,
                                 const double* row_scale, double col_scale, double* partial, int* ctrl, int* dbg, size_t dbg_ld,
                                 cudaStream_t st, const unsigned char* nzA = nullptr, size_t nz_pitch = 0);
// nz[rt*pitch + kb] bit t: slice t of (row tile rt, 64-wide k-block kb) of As holds a nonzero digit (all-zero blocks are skipped)
cudaError_t launch_ozaki_mask(const signed char* As, size_t a_pitch, size_t a_slice, int S, int nrt, int kblocks, unsigned char* nz,
                              size_t pitch, cudaStream_t st);
cudaError_t launch_ozaki_slice_x(const double* X, size_t ld, int n_rows, int S, int base254, signed char* Xs, double* row_scale,
                                 unsigned long long* rowmax, cudaStream_t st);
cudaError_t launch_ozaki_diag_scale(const double* A, size_t ld, int n, double* row_scale, cudaStream_t st);
cudaError_t launch_ozaki_slice_lpanel(const double* A, size_t ld, size_t r0, size_t n_rows, size_t width, const double* row_scale, int S,
                                      signed char* Ls, size_t pitch, size_t slice, cudaStream_t st);
cudaError_t launch_ozaki_syrk_update(const signed char* Ls, size_t pitch, size_t slice, int S, size_t r0, size_t n_rows, size_t width,
                                     double* A, size_t ld, const double* row_scale, int* ctrl, cudaStream_t st,
                                     const unsigned char* nz = nullptr, size_t nz_pitch = 0);
cudaError_t launch_ozaki_slice_panel(const double* panel, size_t panel_ld, int q, int n_k, double inv_scale, int S, int base254,
                                     signed char* Ks, size_t k_pitch, size_t q_pad, cudaStream_t st);
// K5 (gpr_append.cu): one slab of k <= 32 appended points at rows [n0, n0+k); ws: append_workspace_doubles(cap).
size_t append_workspace_doubles(size_t cap);
cudaError_t launch_append_slab(const double* xyz, size_t ld, const double* sigma2, int n0, int k, double* L, double* X,
                               double* Dinv, double* ws, size_t cap, const KernParams& kp, int reset_flag,
                               cudaStream_t st);
const int* append_flag_ptr(const double* ws, size_t cap);
cudaError_t launch_identity_rows(double* L, double* X_or_null, size_t ld, int r0, int r1, int row_end, cudaStream_t st);
cudaError_t launch_dinv_from_x(const double* X, size_t ld, int tile0, int ntiles, double* Dinv, cudaStream_t st);
cudaError_t launch_skinny(int mode, const double* A, size_t ld, int rows, int kdim, const double* Bm, double* OUT,
                          cudaStream_t st);
cudaError_t launch_append_panel(const double* xyz, size_t ld, const double* sigma2, int n0, int t0, int k, double* Pn,
                                double* S0, const KernParams& kp, cudaStream_t st);
// Indefinite tail (gpr_tail.cu)
cudaError_t launch_conflict_counts(const double* xyz, size_t ld, const double* sigma2, int n, int* counts,
                                   const KernParams& kp, cudaStream_t st);
cudaError_t launch_tail_cc(const double* xyz, size_t ld, const double* sigma2, int p, int m, int mp, double* C,
                           const KernParams& kp, cudaStream_t st);
size_t tail_gram_part_doubles(int p, int mp);
cudaError_t launch_tail_schur(const double* B, size_t ldr, int p, int mp, const double* C, double* part, double* S,
                              cudaStream_t st);
cudaError_t launch_tail_alpha(const double* B, const double* Z, size_t ldr, int p, int m, int mp, const double* zf,
                              const double* label, const double* Sinv, double* t, double* a2, double* alpha,
                              cudaStream_t st);
cudaError_t launch_tail_var(const double* qx, const double* qy, const double* qz, int q, const double* xyz, size_t ld,
                            int p, int m, int mp, const double* W, size_t ldq, const double* Sinv, double* var,
                            const KernParams& kp, cudaStream_t st);
// Engine self-test (gpr_selftest.cu): C = A * B^T on one 128x128 tile per CTA.
cudaError_t launch_gemm_selftest(const double* A, size_t lda, const double* B, size_t ldb, int b_kmajor, double* C,
                                 size_t ldc, int mt, int nt, int k, cudaStream_t st);
cudaError_t launch_leaf_selftest(double* tile /*128x128 in/out: L*/, double* inv /*128x128 out*/, int* info,
                                 cudaStream_t st, long long* cycles = nullptr);

cudaError_t run_peak_probe(int which, int ctas_per_sm, double* tflops);

}  // namespace gpr

/* gpr_c_api.h — C-ABI of the B200-native GP-regression core (libgpr_b200.so).
 *
 * This is the drop-in boundary for the hot path of pacman-project/gaussian-object-modelling:
 * everything that the reference's header-only template
 *     /root/reference/include/gp_regression/gp_regressor.hpp   (class GPRegressor<CovType>)
 * computes with Eigen on one CPU thread is computed here by hand-written sm_100a CUDA kernels.
 * Plain C types only: SoA double arrays exactly as in gp_regression::Data
 * (gp_regressor.hpp:49-55), sizes, opaque handles.  There is no CPU fallback: every entry point
 * that computes fails with GPR_ERR_CUDA when no CUDA device is usable.
 *
 * The C++ headers in include/gp_regression/ are thin shims over this file and keep the reference's
 * names, signatures and exception messages; INTEGRATION.md shows the binding.
 *
 * Matrices crossing this boundary are column-major with the stated leading dimension, like the
 * Eigen::MatrixXd outputs of the reference (q x 3: all x, then all y, then all z).
 */
#ifndef GPR_C_API_H
#define GPR_C_API_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpr_ctx gpr_ctx;
typedef struct gpr_model gpr_model;

/* Status codes.  GPR_ERR_NOT_SPD: the covariance matrix is not positive definite (for the
 * thin-plate kernel: R smaller than the largest pairwise distance); the 1-based index of
 * the failing pivot is returned by gpr_last_pivot(). */
enum { GPR_OK = 0, GPR_ERR_INVALID = 1, GPR_ERR_NOT_SPD = 2, GPR_ERR_CUDA = 3, GPR_ERR_OOM = 4 };

/* Covariance function = the reference's kernel classes:
 *   kind 0  ThinPlate(R = p0)               kernels/thin_plate.hpp:12-38
 *   kind 1  Gaussian(sigma = p0, length = p1)   kernels/gaussian.hpp:15-50   (sigma^2 exp(-d/length^2))
 *   kind 2  Laplace(sigma = p0, length = p1)    kernels/laplace.hpp:37-70    (2 sigma exp(-d/length)) */
typedef struct { int kind; double p0; double p1; } gpr_kernel_t;

/* Per-phase device times of the last fit / predict on this context, CUDA-event measured (ms).
 * Replaces the std::chrono prints of the reference's caller (src/gp_node.cpp:894, :923-925). */
typedef struct {
    double cov_ms, chol_ms, solve_ms, normals_ms, fit_total_ms;
    double linv_ms;                 /* one-time L^-1 for the variance path */
    double predict_mean_ms, predict_var_ms, predict_total_ms;
    double h2d_ms, d2h_ms;
    double append_ms;               /* last gpr_append: device time of the incremental update incl. the alpha re-solve */
    double ozaki_ms;                /* last predict on the primary device: time inside the INT8 tensor-core variance kernel (0 if another form ran) */
    double ozaki_slices;            /* ... and the number of int8 slices per operand it used (6: base-254 digits, 7 or more: escalated / base 128) */
    double ozaki_issued_fraction;   /* ... and the share of its slice-pair MMAs actually issued (digit slices of L^-1 that are all zero in a
                                       (128-row, 64-k) block are skipped) */
    double fit_int8_slices;         /* last fit: digit slices of the INT8-assisted factorisation (0: the all-FP64 tile Cholesky ran) */
} gpr_timings;

/* ---- context ------------------------------------------------------------------------------ */
/* devices == NULL or ndev <= 0: device 0 only.  With several devices, gpr_fit runs on devices[0]
 * and gpr_predict shards the queries over all of them. */
int gpr_ctx_create(const int* devices, int ndev, gpr_ctx** out);
int gpr_ctx_destroy(gpr_ctx* ctx);
int gpr_ctx_num_devices(const gpr_ctx* ctx);
const char* gpr_last_error(void);          /* thread-local message of the last failing call */
long long gpr_last_pivot(void);            /* thread-local, valid after GPR_ERR_NOT_SPD */
int gpr_last_timings(const gpr_ctx* ctx, gpr_timings* out);

/* ---- fit: GPRegressor::create<withNormals>  (gp_regressor.hpp:110-182) ---------------------- */
/* x,y,z,label: n host doubles each; sigma2: n host doubles or NULL (Data::sigma2 empty, :154).
 * The factorisation of models with n >= 8192 (GPR_FIT_INT8_MIN_N) runs INT8-assisted: panels of 16 tile columns on the FP64
 * tensor pipe, everything left of a panel as exact int8 digit-slice products on tcgen05 (7 base-254 digits: the dropped part
 * is below the FP64 rounding of the same sums); GPR_FIT_MODE=fp64 forces the all-FP64 tile Cholesky.  gpr_timings.fit_int8_slices
 * says which ran.  A matrix that turns out not to be positive definite is always re-factorised (and reported) in FP64. */
int gpr_fit(gpr_ctx* ctx, const double* x, const double* y, const double* z, const double* label,
            const double* sigma2_or_null, size_t n, gpr_kernel_t kernel, int with_normals, gpr_model** out);
int gpr_model_destroy(gpr_model* m);
size_t gpr_model_size(const gpr_model* m);                 /* n */
/* Indefinite covariance (the node's ThinPlate(2.0) + external sphere setting): when the Cholesky
 * factorisation meets a non-positive pivot, the offending points (at most 256; found by counting indefinite 2x2
 * minors, wherever they sit in the training set) are moved to the end of an internal order and eliminated as
 * one dense pivot block (block L D L^T, csrc/gpr_tail.cu), and the fit succeeds; this returns how many points
 * are in that block (0 for a positive definite matrix).  Outputs stay in the caller's point order.
 * GPR_NO_TAIL=1 disables it (gpr_fit then returns GPR_ERR_NOT_SPD as for any other indefinite matrix). */
size_t gpr_model_tail_size(const gpr_model* m);
/* Model fields the reference exposes (gp_regressor.hpp:71-87): alpha[n], R, N (n x 3 column-major,
 * only when fitted with normals).  Any output pointer may be NULL. */
int gpr_model_get(const gpr_model* m, double* alpha, double* R, double* normals_or_null);
/* Debug / parity access: the assembled covariance is not kept; the lower Cholesky factor is.
 * L: n x n column-major, strict upper triangle zeroed.  For a model with an indefinite tail block
 * (gpr_model_tail_size() > 0) only the leading (n - tail) x (n - tail) block is a Cholesky factor, in the
 * library's internal point order. */
int gpr_model_get_factor(const gpr_model* m, double* L);

/* ---- predict: the four GPRegressor::evaluate overloads (gp_regressor.hpp:194,:222,:282,:332) - */
/* qx,qy,qz: q host doubles.  f: q.  var: q or NULL.  grad: q x 3 column-major (leading dimension q),
 * un-normalised as in the reference (:247-250), or NULL.  tx, ty: q x 3 tangent basis (:204-211), both
 * or neither; they require grad.  Thread-safe on one model: concurrent calls use separate streams and
 * workspaces (the reference is called from hundreds of threads, src/gp_node.cpp:1027-1038). */
int gpr_predict(gpr_ctx* ctx, gpr_model* m, const double* qx, const double* qy, const double* qz, size_t q,
                double* f, double* var_or_null, double* grad_or_null, double* tx_or_null, double* ty_or_null);
/* Same, all pointers in device memory of the model's primary device (devices[0]); no host copies.
 * Runs on an internal stream and returns after it has drained. */
int gpr_predict_device(gpr_ctx* ctx, gpr_model* m, const double* d_qx, const double* d_qy, const double* d_qz,
                       size_t q, double* d_f, double* d_var_or_null, double* d_grad_or_null);
/* Batched iso-surface sampling — the step right above the path in the reference's node
 * (fakeDeterministicSampling / samplePoint, src/gp_node.cpp:998-1100: one std::thread and one
 * evaluate(q = 1) per lattice point, keep the point if |f| <= 0.01, its variance is the intensity).
 * Lattice: every axis takes the values lo, lo+step, ... <= hi accumulated like the node's loops
 * (x outermost, z innermost).  The mean is evaluated for the whole lattice on the device; the variance
 * only for the points that are kept.  Outputs (host, each `capacity` doubles or NULL) are in lattice order;
 * *count receives the number of lattice points with |f| <= tol (it may exceed capacity: then only the first
 * `capacity` are written). */
int gpr_sample_isosurface(gpr_ctx* ctx, gpr_model* m, double lo, double hi, double step, double tol, size_t capacity,
                          double* x, double* y, double* z, double* f, double* var_or_null, size_t* count);
/* Batched counterpart of the node's marchingSampling / marchingCubes (src/gp_node.cpp:1103-1292; called with leaf 0.06,
 * pass 0.02 at :258): a flood fill over cubes of edge `leaf` that follows the iso-surface — start at the first point of the
 * lattice lo, lo+step, ... <= hi (x outermost; the reference: -1.1 .. 1.1 step 0.1) with |f| <= tol, sample every cube on
 * its (round(leaf/pass) + 1)^3 lattice, keep the samples with |f| <= tol (the variance is their intensity) and expand across
 * every face that a kept sample touches.  The reference runs one std::thread per cube and one evaluate(q = 1) per sample;
 * here every wave of the flood fill is one batched evaluation.  Same float (pcl::PointXYZ) coordinate arithmetic.  Samples
 * on a face shared by two cubes are returned once (the reference removes them afterwards with a PCL voxel filter, which
 * also averages near-duplicates: that PCL post-processing is not reproduced).  Outputs (host, `capacity` each or NULL) in
 * discovery order (waves; cubes of a wave by index); *count = kept samples (may exceed capacity), *cubes = cubes visited.
 * GPR_ERR_INVALID "No starting point found. Relax grid pass." as the reference prints (:1153). */
int gpr_sample_marching(gpr_ctx* ctx, gpr_model* m, double lo, double hi, double step, float leaf, float pass, double tol,
                        size_t capacity, double* x, double* y, double* z, double* f, double* var_or_null, size_t* count,
                        size_t* cubes_or_null);
/* Batched projection onto the iso-surface f = 0 — AtlasBase::project (include/atlas/atlas.hpp:201-276: fixed-step
 * gradient descent, two evaluate(q = 1) calls per iteration, up to max_iter = 500 iterations per point).
 * Same update rule (x -= step_mul * f(x) * g, g = last accepted un-normalised gradient, first one given by the
 * caller) and the same stopping criteria (|f| < f_tol, |f_new - f_old| < improve_tol, max_iter); the whole
 * iteration of every point runs on the device in ONE launch.  x..nz: count host doubles each; ox, oy, oz: count
 * each; status (count ints or NULL): iterations used when a tolerance was met, -max_iter when the budget ran
 * out.  Returns GPR_ERR_INVALID "f is nan or inf" if that happened for any point (the reference throws). */
int gpr_project(gpr_ctx* ctx, gpr_model* m, const double* x, const double* y, const double* z, const double* nx,
                const double* ny, const double* nz, size_t count, double f_tol, double improve_tol, unsigned max_iter,
                double step_mul, double* ox, double* oy, double* oz, int* status_or_null);
/* Batched AtlasVariance::sampleOnChart (include/atlas/atlas_variance.hpp:147-219: per chart, ceil(|disc_samples_factor| R)
 * uniform annulus samples on the tangent disc, one evaluate(f, v) each, then sorted by decreasing variance) for any number
 * of charts in one call: the samples are generated on the device, evaluated as ONE batch, and ranked per chart on the device.
 *   frames : n_charts x 13 doubles, per chart: centre C (3), normal N (3), tangent basis Tx (3), Ty (3), radius R
 *            (Chart::getCenter / getNormal / getTanBasisOne / getTanBasisTwo / getRadius, include/atlas/atlas.hpp:17-106)
 *   counts : samples per chart;  total = their sum
 *   r, th  : `total` uniform variates each, r in [0.8, 1], th in [0, 2 pi), in chart order — the reference draws them from a
 *            process-global mt19937_64 seeded by std::random_device (include/random_generation.hpp:9-27), i.e. NOT reproducibly;
 *            pass both for a deterministic result, or both NULL to have them drawn from mt19937_64(seed) with the reference's
 *            distributions and call order (r, then th, per sample)
 *   sx, sy, sz, f, v (host, `total` each, any may be NULL): sample points (Chart::samples rows), their mean and variance
 *   order  (host, `total`, may be NULL): for chart c with offset o_c, order[o_c + k] = index within the chart of the sample
 *            with the k-th largest variance (Chart::vars_ids after the sort, :214-218; equal variances by index)
 * Returns GPR_ERR_INVALID "v is nan or inf" if any f or v is not finite (the reference throws, :202-211). */
int gpr_sample_chart(gpr_ctx* ctx, gpr_model* m, const double* frames, const size_t* counts, size_t n_charts,
                     const double* r_or_null, const double* th_or_null, unsigned long long seed, double* sx, double* sy,
                     double* sz, double* f, double* v, size_t* order);
/* Builds L^-1 now.  It is otherwise built by the first call that needs it.  The variance (n^2 flop per query) has three
 * forms, chosen per call (GPR_VAR_MODE=ozaki|product|trsm forces one):
 *   >= 16384 queries (GPR_OZAKI_MIN_Q): product with L^-1 on the INT8 tensor cores (tcgen05 kind::i8 + TMEM + TMA),
 *      FP64-equivalent by Ozaki slicing (6 slices of base-254 digits; rows longer than 22016 drain the int32 accumulators in chunks:
 *      ~1e-9 of the variance; GPR_OZAKI_SLICES / GPR_OZAKI_BASE override), ~2.9x the FP64 tensor-pipe rate; needs L^-1 and its
 *      int8 slices (built once per model); every call re-computes its first query tile on the FP64 tensor pipe and, if the
 *      two differ by > 1e-8, adds a slice for this model (remembered) or, failing that, falls back to the FP64 product form;
 *      once a model holds its slices every batch of more than 8 queries uses them (results do not depend on call sizes);
 *   otherwise, if L^-1 is resident (or < 4096 queries, GPR_TRSM_MIN_Q: the fused single-query kernel needs it anyway, and
 *      so does gpr_append): product with L^-1 on the FP64 tensor pipe;
 *   otherwise: forward substitution over L (no inverse at all, one n x n matrix per model).
 * Models with an indefinite tail block always use the FP64 product form. */
int gpr_model_prepare_variance(gpr_ctx* ctx, gpr_model* m);
/* X = K^-1 B for nrhs right-hand sides (B, X: n x nrhs column-major, host) through the resident factor — the
 * counterpart of the reference's public Model::cholesker.solve(b) (gp_regressor.hpp:81, :163).  Positive definite
 * models only (GPR_ERR_INVALID for a model with an indefinite tail block). */
int gpr_model_solve(gpr_ctx* ctx, gpr_model* m, const double* B, size_t nrhs, double* X);

/* ---- update: GPRegressor::update<withNormals>  (gp_regressor.hpp:367-479) --------------------- */
/* Appends k points; R and the normals are not refreshed, as in the reference (:454-455, :462-477).
 * The reference re-factorises from scratch (:457-459).  Here small batches (k <= 256 and 8k <= n) take the
 * incremental path: the new rows of the Cholesky factor and of its inverse are appended (slabs of 32
 * points, two bandwidth-bound products against L^-1 each) and alpha is re-solved through L^-1 with one step of
 * iterative refinement (GPR_APPEND_TRSV=1: by the triangular solves over L); larger batches refit.
 * A model with an indefinite tail block (the node's real setting, whose cb_update refits per touch, src/gp_node.cpp:652-763)
 * is updated incrementally too: the new points join the positive definite leading block (rows of L and L^-1 appended,
 * the tail points moved behind them in the internal order) and the trailing block is eliminated again against the
 * extended L^-1; if a new point does not fit the leading block (non-positive pivot) the call falls back to the refit.
 * On GPR_ERR_NOT_SPD the model is left as it was before the call.  GPR_APPEND_REFIT=1 forces the refit. */
int gpr_append(gpr_ctx* ctx, gpr_model* m, const double* x, const double* y, const double* z, const double* label,
               const double* sigma2_or_null, size_t k);
/* Pre-allocates room for `capacity` points so that later appends do not reallocate (growth is otherwise
 * geometric, x1.25).  Invalidates pointers returned earlier by gpr_model_state_get.  No effect on a model with an
 * indefinite tail block (it grows on demand when it is appended to). */
int gpr_model_reserve(gpr_ctx* ctx, gpr_model* m, size_t capacity);

/* ---- export / import (the reference has no persistence for the GP) ----------------- */
/* Writes the training set, kernel, R, alpha, normals and — if with_factor != 0 and the matrix was positive
 * definite — the lower Cholesky factor (n(n+1)/2 doubles) to a binary file.  gpr_model_load restores the model
 * on ctx's primary device: with a stored factor without refactorising (only the 128x128 diagonal inverses are
 * rebuilt; L^-1 is rebuilt on the first variance request), otherwise by refitting the stored training set. */
int gpr_model_save(gpr_ctx* ctx, gpr_model* m, const char* path, int with_factor);
int gpr_model_load(gpr_ctx* ctx, const char* path, gpr_model** out);

/* ---- point-cloud input ----------------------------------------------------------------- */
/* Minimal PCD v0.7 reader — what the reference obtains from PCL (pcl::io::loadPCDFile, src/gp_node.cpp:557): the x, y, z
 * fields (4-byte floats, widened to double) of an ascii / binary / binary_compressed (LZF) file, as three malloc'ed arrays
 * of *n doubles that the caller releases with gpr_free.  Other fields (rgba, normals) are skipped.  Host only. */
int gpr_pcd_read_xyz(const char* path, double** x, double** y, double** z, size_t* n);
void gpr_free(void* p);

/* ---- replication across processes (one process per GPU, launched by bench.py) ----------------- */
/* The fitted state that predict needs, as raw device pointers on the primary device, so that the
 * launcher can broadcast it with NCCL into a model created by gpr_model_create_replica on another rank.
 * padded_n = 128*ceil(n/128); xyz: 3*padded_n (x | y | z); alpha: padded_n; linv: padded_n^2 (or NULL
 * until the variance path was prepared). */
typedef struct {
    size_t n, padded_n;
    size_t ld;                   /* leading dimension of xyz (x | y | z at multiples of ld), alpha and linv;
                                    == padded_n unless the model grew through gpr_append / gpr_model_reserve */
    gpr_kernel_t kernel;
    double R;
    double* xyz; double* alpha; double* linv;
    /* indefinite tail block (0 / NULL for a positive definite matrix): tail_z is (tail_pad/32) slabs of ld x 32
     * doubles, tail_sinv tail_pad x tail_pad; linv / the factor then describe the leading n - n_tail points */
    size_t n_tail, tail_pad;
    double* tail_z; double* tail_sinv;
    /* the Cholesky factor itself (ld x ld, lower 128x128 tiles) and the inverses of its nb = padded_n/128 diagonal
     * blocks (nb tiles of 128x128, leading dimension 128): what the variance by forward substitution reads.
     * NULL for a model with an indefinite tail block. */
    double* lfac; double* dinv;
} gpr_model_state;
/* with_linv is a bit set here and in the two calls below: 1 = L^-1 (built now if it was not yet), 2 = the factor
 * L + Dinv (nothing to build: what a fit leaves behind). */
int gpr_model_state_get(gpr_ctx* ctx, gpr_model* m, int with_linv, gpr_model_state* out);
/* Allocates an un-fitted model of the given size on ctx's primary device; the caller fills the buffers
 * returned by gpr_model_state_get(replica, ...) (e.g. as the destination of a broadcast).  A replica that holds
 * only the factor (with_linv = 2) computes large-batch variances by forward substitution; batches of <= 8 queries
 * need L^-1 (with_linv & 1). */
int gpr_model_create_replica(gpr_ctx* ctx, size_t n, gpr_kernel_t kernel, double R, int with_linv, gpr_model** out);
/* Same for a model whose last n_tail points (internal order) form the indefinite tail block (gpr_model_tail_size). */
int gpr_model_create_replica_tail(gpr_ctx* ctx, size_t n, size_t n_tail, gpr_kernel_t kernel, double R, int with_linv,
                                  gpr_model** out);
/* Replication fused into the factorisation (one process per GPU on one NVLink box).  Instead of broadcasting the n x n
 * factor after the fit, every rank but the fitting one creates its replica first (gpr_model_create_replica with
 * with_linv = 2) and exports CUDA IPC handles of its factor buffers (128 bytes: L, then Dinv); the launcher gathers them
 * on the fitting rank, which registers them; from then on every gpr_fit on that context whose padded size matches stores
 * each finished 128 x 128 tile of L and Dinv into all registered replicas from inside the Cholesky kernel (posted NVLink
 * peer writes), so the replicas are complete when gpr_fit returns; only {x|y|z, alpha} (32 n bytes) remain to be
 * broadcast.  gpr_ctx_last_fit_published tells whether the last fit published (it does not when the matrix turned out
 * indefinite and the trailing-block path was taken: broadcast L^-1 and the tail block then).  The replicas may be used
 * after the launcher's barrier that follows gpr_fit.  At most 7 peers. */
int gpr_model_ipc_export(gpr_ctx* ctx, gpr_model* replica, void* handles128);
int gpr_ctx_set_fit_peers(gpr_ctx* ctx, const void* handles128_each, int n_peers, size_t n);
int gpr_ctx_clear_fit_peers(gpr_ctx* ctx);
int gpr_ctx_last_fit_published(const gpr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif

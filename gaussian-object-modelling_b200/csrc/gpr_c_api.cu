// gpr_c_api.cu — implementation of include/gpr_c_api.h: contexts, models, the fit / predict / append
// orchestration over the kernels in this directory.  No CPU fallback anywhere: every compute entry
// point needs a CUDA device and reports GPR_ERR_CUDA otherwise.
#include "../../include/gpr_c_api.h"
#include "gpr_selftest.h"
#include "gpr_kernels.h"
#include "gpr_mma.cuh"

#include <algorithm>
#include <array>
#include <set>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <limits>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

using namespace gpr;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local long long g_pivot = 0;

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
void gpr_set_last_error_internal(const char* msg) { g_err = msg ? msg : ""; }      // for gpr_io.cu

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return fail(e__ == cudaErrorMemoryAllocation ? GPR_ERR_OOM : GPR_ERR_CUDA, b__);       \
        }                                                                                          \
    } while (0)

// Debug aid: GPR_POISON=1 fills every fresh device allocation of this file with 0xFF bytes (NaN doubles, huge
// ints), so that any read of memory the library has not written shows up as NaN in the parity tests instead of
// depending on what cudaMalloc happens to return (a fresh process gets zeroed pages, a long-running one does not).
static cudaError_t gpr_malloc_poison(void** p, size_t bytes) {
    static const bool poison = getenv("GPR_POISON") && atoi(getenv("GPR_POISON")) != 0;
    cudaError_t e = (cudaMalloc)(p, bytes);
    if (e == cudaSuccess && poison && bytes) {
        e = cudaMemset(*p, 0xFF, bytes);          // legacy stream: not ordered against the library's non-blocking streams
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    return e;
}
#define cudaMalloc(p, bytes) gpr_malloc_poison((void**)(p), (bytes))

// ------------------------------------------------------------------------------------------------
// context, workspaces
// ------------------------------------------------------------------------------------------------
struct Workspace {
    int dev = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[8] = {};
    double* io = nullptr; size_t io_cap = 0;          // 14 * io_cap doubles: q(3) f var grad(3) tx(3) ty(3)
    double* panel = nullptr; size_t panel_dbl = 0;
    double* partial = nullptr; size_t partial_dbl = 0;
    double* mpart = nullptr; size_t mpart_dbl = 0;     // per-chunk partial sums of the split mean kernel
    double* tailw = nullptr; size_t tailw_dbl = 0;     // W = Z^T k_1 of the indefinite-tail correction
    double* hio = nullptr; double* hio_dev = nullptr;  // pinned host buffer mapped into the device (q <= 8 path)
    double* small = nullptr; size_t small_dbl = 0; size_t small_N = 0;   // scratch + tickets of the fused small-batch kernel
    signed char* oz_ks = nullptr; size_t oz_ks_bytes = 0;                // int8 slices of the K* panel (gpr_ozaki.cu)
    int* oz_ctrl = nullptr;
};

struct DeviceCtx {
    int dev = 0;
    int num_sms = 148;
    std::mutex mu;
    std::vector<Workspace*> free_ws;
};

struct gpr_ctx {
    std::vector<DeviceCtx*> devs;
    gpr_timings timings;
    std::mutex tmu;
    size_t query_tile = 0;     // queries per variance batch (multiple of 128)
    int chol_serial = 0;
    int refine_steps = 1;      // iterative-refinement steps of alpha after the triangular solves
    // The n x n buffers (L, L^-1) of the most recently destroyed models, kept for the next fit of the same size
    // (a refit per touch, src/gp_node.cpp:750-751, would otherwise pay cudaMalloc/cudaFree of gigabytes each time).
    struct Big { void* p; size_t bytes; int dev; };
    std::vector<Big> big_cache;
    std::mutex cmu;
    // Replicas on other GPUs (other processes: CUDA IPC mappings) that the next fits publish their factor into while the
    // Cholesky kernel runs (gpr_ctx_set_fit_peers).  peer_N: padded size the mappings were made for.
    CholPeers peers{};
    std::vector<void*> peer_maps;          // what cudaIpcOpenMemHandle returned (closed by gpr_ctx_clear_fit_peers)
    size_t peer_N = 0;
    bool last_fit_published = false;
};

static std::mutex g_live_mu;
static std::vector<gpr_ctx*> g_live_ctx;                 // contexts that have not been destroyed (a model may outlive its context)
constexpr size_t BIG_MIN = (size_t)64 << 20, BIG_MAX_TOTAL = (size_t)9 << 30;

static bool big_cache_off() { static const bool off = getenv("GPR_NO_CACHE") && atoi(getenv("GPR_NO_CACHE")) != 0; return off; }

static void* big_take(gpr_ctx* ctx, int dev, size_t bytes) {
    if (big_cache_off()) return nullptr;
    std::lock_guard<std::mutex> lk(ctx->cmu);
    for (size_t i = 0; i < ctx->big_cache.size(); ++i)
        if (ctx->big_cache[i].dev == dev && ctx->big_cache[i].bytes == bytes) {
            void* p = ctx->big_cache[i].p;
            ctx->big_cache.erase(ctx->big_cache.begin() + (std::ptrdiff_t)i);
            return p;
        }
    return nullptr;
}
// cudaMalloc through the cache.  The buffer is NOT cleared.
static cudaError_t big_alloc(gpr_ctx* ctx, int dev, void** p, size_t bytes) {
    *p = bytes >= BIG_MIN ? big_take(ctx, dev, bytes) : nullptr;
    if (*p) {
        static const bool poison = getenv("GPR_POISON") && atoi(getenv("GPR_POISON")) != 0;
        if (!poison) return cudaSuccess;
        cudaError_t e = cudaMemset(*p, 0xFF, bytes);
        return e == cudaSuccess ? cudaDeviceSynchronize() : e;
    }
    return cudaMalloc(p, bytes);
}
// Returns the buffer to the cache of its context (if that context is still alive and the cache has room), else frees it.
// The caller guarantees that no work is pending on the buffer.  Current device must be `dev`.
static void big_free(gpr_ctx* ctx, int dev, void* p, size_t bytes) {
    if (!p) return;
    if (!big_cache_off() && bytes >= BIG_MIN && bytes <= BIG_MAX_TOTAL / 2) {
        std::lock_guard<std::mutex> lk(g_live_mu);
        if (std::find(g_live_ctx.begin(), g_live_ctx.end(), ctx) != g_live_ctx.end()) {
            std::lock_guard<std::mutex> lk2(ctx->cmu);
            size_t total = bytes;
            for (auto& b : ctx->big_cache) total += b.bytes;
            while (!ctx->big_cache.empty() && (total > BIG_MAX_TOTAL || ctx->big_cache.size() >= 4)) {
                total -= ctx->big_cache.front().bytes;
                cudaSetDevice(ctx->big_cache.front().dev);
                cudaFree(ctx->big_cache.front().p);
                ctx->big_cache.erase(ctx->big_cache.begin());
            }
            cudaSetDevice(dev);
            ctx->big_cache.push_back(gpr_ctx::Big{p, bytes, dev});
            return;
        }
    }
    cudaFree(p);
}

static int ws_acquire(DeviceCtx* dc, Workspace** out) {
    {
        std::lock_guard<std::mutex> lk(dc->mu);
        if (!dc->free_ws.empty()) { *out = dc->free_ws.back(); dc->free_ws.pop_back(); return GPR_OK; }
    }
    Workspace* ws = new Workspace();
    ws->dev = dc->dev;
    CU(cudaSetDevice(dc->dev));
    CU(cudaStreamCreateWithFlags(&ws->st, cudaStreamNonBlocking));
    for (auto& e : ws->ev) CU(cudaEventCreate(&e));
    *out = ws;
    return GPR_OK;
}
static void ws_release(DeviceCtx* dc, Workspace* ws) {
    std::lock_guard<std::mutex> lk(dc->mu);
    dc->free_ws.push_back(ws);
}
static int ws_reserve(double** p, size_t* cap, size_t need) {
    if (*cap >= need) return GPR_OK;
    if (*p) CU(cudaFree(*p));
    *p = nullptr; *cap = 0;
    CU(cudaMalloc((void**)p, need * sizeof(double)));
    *cap = need;
    return GPR_OK;
}

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
struct ModelDev {            // what predict needs, per device
    int dev = 0;
    double* xyz = nullptr;   // 3N: x | y | z
    double* alpha = nullptr; // N
    double* linv = nullptr;  // N x N, lower tiles
    double* tZ = nullptr;    // indefinite tail (gpr_tail.cu): Z = A^-1 P in 32-column slabs, S^-1
    double* tSinv = nullptr;
    // Cholesky factor + inverses of its diagonal blocks, for the variance by forward substitution (gpr_var.cu:
    // var_trsm_kernel).  On the primary device these alias gpr_model::L / Dinv (own_fac false); copies on the other
    // devices of the context and in cross-process replicas are owned.
    double* lfac = nullptr; double* dinv = nullptr;
    bool have = false, have_linv = false, have_tail = false, have_fac = false, own_fac = false;
    // int8 slices of X = L^-1 and its per-row power-of-two scales, for the variance on the INT8 tensor cores (gpr_ozaki.cu)
    signed char* oz_xs = nullptr; double* oz_scale = nullptr; int oz_S = 0, oz_base = 0; size_t oz_ld = 0, oz_n = 0;   // oz_n: model size the slices were cut for
    unsigned char* oz_nz = nullptr; double oz_exec = 1.0;      // nonzero-slice map per (row tile, k-block); fraction of the MMAs that are issued
};

struct gpr_model {
    gpr_ctx* ctx = nullptr;
    size_t n = 0, N = 0;       // real points; active padded size 128*ceil(n/128)
    size_t cap = 0;            // leading dimension of every per-point buffer and of L / L^-1 (>= N, multiple of 128)
    int nb = 0;
    double* aws = nullptr; size_t aws_dbl = 0;   // append workspace (primary device)
    // Indefinite tail (gpr_tail.cu): the last n_tail points are eliminated as one dense pivot block; L / L^-1
    // / nb then describe the leading n_spd = n - n_tail points only.  n_tail == 0 for SPD matrices.
    size_t n_spd = 0, n_tail = 0;
    std::vector<size_t> perm;                    // internal index -> caller's index (empty = identity)
    int mp = 0;                                  // n_tail rounded up to a multiple of 32
    double* tB = nullptr; double* tmisc = nullptr;   // B = X P (slabs); C | S | t | a2 | gram partials | panel tmp
    gpr_kernel_t kernel{};
    KernParams kp{};
    double R = 0.0, k0 = 0.0;
    bool has_s2 = false, with_normals = false, replica = false;
    // factor state on the primary device
    double* label = nullptr; double* s2 = nullptr; double* zfwd = nullptr;
    double* L = nullptr; double* Dinv = nullptr; int* scratch = nullptr;
    std::vector<ModelDev> devs;
    std::vector<double> hx, hy, hz, hlabel, hs2, h_alpha, h_normals;   // h_normals: n_normals x 3 column-major
    size_t n_normals = 0;
    int fit_int8_slices = 0;         // digit slices of the INT8-assisted factorisation that produced L (0: all-FP64)
    bool oz_disabled = false;        // the INT8 tensor-core variance failed its FP64 spot check on this model: FP64 paths only
    int oz_bump = 0;                 // extra slices this model needs beyond the default (raised by a failed spot check)
    std::vector<std::pair<int, void*>> retired;   // (device, buffer): slice buffers replaced while concurrent predict calls may
                                                  // still read them (slice escalation); freed with the model
    std::mutex mu;
    // Micro-batcher of the callers' q = 1 pattern (hundreds of concurrent threads with one query each on one shared
    // model, src/gp_node.cpp:1027-1038): concurrent small requests are combined into one batched launch.
    struct SmallReq {
        const double *qx, *qy, *qz; size_t q;
        double *f, *var, *grad, *tx, *ty; size_t out_ld;
        int rc = 0; std::string err; bool done = false;
        std::condition_variable cv;        // one per request: a finished launch wakes exactly the threads it served
    };
    std::mutex bmu;
    std::vector<SmallReq*> pending;
    bool leader = false;
};

static KernParams make_kp(gpr_kernel_t k) {
    KernParams kp;
    kp.kind = k.kind; kp.p0 = k.p0; kp.p1 = k.p1;
    kp.R3 = k.p0 * k.p0 * k.p0;                                    // thin_plate.hpp:31
    if (k.kind == 1) { kp.amp = k.p0 * k.p0; kp.inv = 1.0 / (k.p1 * k.p1); }      // gaussian.hpp:40-41
    else if (k.kind == 2) { kp.amp = 2 * k.p0; kp.inv = 1.0 / k.p1; }             // laplace.hpp:40, :62
    else { kp.amp = 0; kp.inv = 0; }
    return kp;
}
static double kernel_at_zero(const KernParams& kp) { return kp.kind == 0 ? kp.R3 : kp.amp; }

static void free_factor(gpr_model* m) {
    for (auto& r : m->retired) { cudaSetDevice(r.first); cudaFree(r.second); }
    m->retired.clear();
    if (m->devs.empty()) return;
    cudaSetDevice(m->devs[0].dev);
    cudaFree(m->label); cudaFree(m->s2); cudaFree(m->zfwd); cudaFree(m->Dinv); cudaFree(m->scratch);
    big_free(m->ctx, m->devs[0].dev, m->L, m->cap * m->cap * sizeof(double));
    cudaFree(m->aws); cudaFree(m->tB); cudaFree(m->tmisc);
    m->label = m->s2 = m->zfwd = m->L = m->Dinv = m->aws = m->tB = m->tmisc = nullptr; m->scratch = nullptr; m->aws_dbl = 0;
    m->n_tail = 0; m->mp = 0;
    for (auto& d : m->devs) {
        cudaSetDevice(d.dev);
        cudaFree(d.xyz); cudaFree(d.alpha); cudaFree(d.tZ); cudaFree(d.tSinv);
        big_free(m->ctx, d.dev, d.linv, m->cap * m->cap * sizeof(double));
        if (d.own_fac) { big_free(m->ctx, d.dev, d.lfac, m->cap * m->cap * sizeof(double)); cudaFree(d.dinv); }
        cudaFree(d.oz_xs); cudaFree(d.oz_scale); cudaFree(d.oz_nz);
        d.oz_xs = nullptr; d.oz_scale = nullptr; d.oz_nz = nullptr; d.oz_S = 0; d.oz_ld = 0;
        d.xyz = d.alpha = d.linv = d.tZ = d.tSinv = d.lfac = d.dinv = nullptr;
        d.have = d.have_linv = d.have_tail = d.have_fac = d.own_fac = false;
    }
}

// Inverse of the symmetric (possibly indefinite) mp x mp matrix S by Gauss-Jordan elimination with partial
// pivoting, on the host (mp <= 256).  Returns false if S is numerically singular.
static bool invert_small(std::vector<double>& S, int mp, std::vector<double>& inv) {
    inv.assign((size_t)mp * mp, 0.0);
    for (int i = 0; i < mp; ++i) inv[(size_t)i * mp + i] = 1.0;
    double scale = 0.0;
    for (double v : S) scale = std::max(scale, std::fabs(v));
    for (int c = 0; c < mp; ++c) {
        int piv = c;
        for (int r = c + 1; r < mp; ++r) if (std::fabs(S[(size_t)r * mp + c]) > std::fabs(S[(size_t)piv * mp + c])) piv = r;
        if (!(std::fabs(S[(size_t)piv * mp + c]) > 1e-14 * scale)) return false;
        if (piv != c) for (int k = 0; k < mp; ++k) { std::swap(S[(size_t)c * mp + k], S[(size_t)piv * mp + k]); std::swap(inv[(size_t)c * mp + k], inv[(size_t)piv * mp + k]); }
        const double d = 1.0 / S[(size_t)c * mp + c];
        for (int k = 0; k < mp; ++k) { S[(size_t)c * mp + k] *= d; inv[(size_t)c * mp + k] *= d; }
        for (int r = 0; r < mp; ++r) {
            if (r == c) continue;
            const double f = S[(size_t)r * mp + c];
            if (f == 0.0) continue;
            for (int k = 0; k < mp; ++k) { S[(size_t)r * mp + k] -= f * S[(size_t)c * mp + k]; inv[(size_t)r * mp + k] -= f * inv[(size_t)c * mp + k]; }
        }
    }
    for (int a = 0; a < mp; ++a)                       // symmetrise
        for (int b = 0; b < a; ++b) { const double v = 0.5 * (inv[(size_t)a * mp + b] + inv[(size_t)b * mp + a]); inv[(size_t)a * mp + b] = inv[(size_t)b * mp + a] = v; }
    return true;
}

constexpr size_t MAX_TAIL = 256;

// Elimination of the trailing pivot block of an indefinite matrix (gpr_tail.cu), given the Cholesky factor of the leading
// n_spd points in m->L: (re)builds B, Z, S^-1 and alpha.  build_linv: also form X = L^-1 of the leading block (the fit);
// an incremental append has already brought X up to date.  Frees and reallocates the tail buffers (their leading
// dimension is the model capacity).
static int tail_eliminate(gpr_model* m, DeviceCtx* dc, Workspace* ws, bool build_linv) {
    gpr_ctx* ctx = m->ctx;
    ModelDev& md = m->devs[0];
    cudaStream_t st = ws->st;
    const size_t N = m->cap;
    cudaFree(m->tB); cudaFree(md.tZ); cudaFree(md.tSinv); cudaFree(m->tmisc);
    m->tB = md.tZ = md.tSinv = m->tmisc = nullptr;
    const size_t p = m->n_spd, mt = m->n_tail;
    const int nslab = (int)((mt + 31) / 32), mp = 32 * nslab;
    m->mp = mp;
    const size_t part_dbl = tail_gram_part_doubles((int)p, mp);
    // tmisc: C | S | t | a2 | gram partials | panel tmp (N x 32) | rhs (N) | S0 dummy (1024)
    const size_t misc = 2 * (size_t)mp * mp + 2 * mp + part_dbl + 32 * N + N + 1024;
    CU(cudaMalloc((void**)&m->tB, (size_t)nslab * N * 32 * sizeof(double)));
    CU(cudaMalloc((void**)&md.tZ, (size_t)nslab * N * 32 * sizeof(double)));
    CU(cudaMalloc((void**)&md.tSinv, (size_t)mp * mp * sizeof(double)));
    CU(cudaMalloc((void**)&m->tmisc, misc * sizeof(double)));
    if (!md.linv) CU(big_alloc(ctx, dc->dev, (void**)&md.linv, N * N * sizeof(double)));
    double* tC = m->tmisc; double* tS = tC + (size_t)mp * mp; double* tt = tS + (size_t)mp * mp; double* ta2 = tt + mp;
    double* tpart = ta2 + mp; double* tPn = tpart + part_dbl; double* trhs = tPn + 32 * N; double* tS0 = trhs + N;
    // alpha_1 part 1: z_f = X y_1, z_1 = A^-1 y_1 (the tail labels must not enter the leading solve)
    CU(cudaMemsetAsync(trhs, 0, N * sizeof(double), st));
    CU(cudaMemcpyAsync(trhs, m->label, p * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CU(launch_trsv(0, m->L, N, m->nb, m->Dinv, trhs, m->zfwd, m->scratch, dc->num_sms, st));
    CU(launch_trsv(1, m->L, N, m->nb, m->Dinv, m->zfwd, md.alpha, m->scratch, dc->num_sms, st));
    // X = L^-1 of the leading block, then B = X P and Z = X^T B slab by slab
    if (build_linv) CU(launch_linv(m->L, md.linv, N, m->nb, m->Dinv, m->scratch, dc->num_sms, st));
    for (int s = 0; s < nslab; ++s) {
        const int kk = (int)std::min<size_t>(32, mt - 32 * (size_t)s);
        double* Bs = m->tB + (size_t)s * N * 32;
        double* Zs = md.tZ + (size_t)s * N * 32;
        CU(launch_append_panel(md.xyz, N, m->s2, (int)p, (int)(p + 32 * (size_t)s), kk, tPn, tS0, m->kp, st));
        CU(launch_skinny(0, md.linv, N, (int)p, (int)p, tPn, Bs, st));
        CU(launch_skinny(1, md.linv, N, (int)p, (int)p, Bs, Zs, st));
    }
    CU(launch_tail_cc(md.xyz, N, m->s2, (int)p, (int)mt, mp, tC, m->kp, st));
    CU(launch_tail_schur(m->tB, N, (int)p, mp, tC, tpart, tS, st));
    std::vector<double> hS((size_t)mp * mp), hSinv;
    CU(cudaMemcpyAsync(hS.data(), tS, hS.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    int lflags[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(lflags, m->scratch, sizeof(lflags), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (lflags[2] != 0) return fail(GPR_ERR_CUDA, "L^-1 kernel aborted (dependency wait timed out)");
    if (!invert_small(hS, mp, hSinv)) {
        g_pivot = (long long)p + 1;
        return fail(GPR_ERR_NOT_SPD, "covariance matrix is singular: the Schur complement of the trailing pivot block cannot be inverted");
    }
    CU(cudaMemcpyAsync(md.tSinv, hSinv.data(), hSinv.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(launch_tail_alpha(m->tB, md.tZ, N, (int)p, (int)mt, mp, m->zfwd, m->label, md.tSinv, tt, ta2, md.alpha, st));
    CU(cudaStreamSynchronize(st));      // hSinv must outlive the copy
    md.have_linv = true; md.have_tail = true;
    return GPR_OK;
}

// ------------------------------------------------------------------------------------------------
// fit
// ------------------------------------------------------------------------------------------------
static float ev_ms(cudaEvent_t a, cudaEvent_t b) { float t = 0; cudaEventElapsedTime(&t, a, b); return t; }

// (Re)fit from the host copies held in the model.  keep_R: the reference's update() does not refresh R.
static int fit_from_host(gpr_model* m, bool keep_R) {
    gpr_ctx* ctx = m->ctx;
    DeviceCtx* dc = ctx->devs[0];
    CU(cudaSetDevice(dc->dev));
    free_factor(m);
    const size_t n = m->hx.size();
    const size_t N = (n + TB - 1) / TB * TB;
    const int nb = (int)(N / TB);
    m->n = n; m->N = N; m->nb = nb; m->cap = N; m->n_spd = n; m->n_tail = 0; m->mp = 0;
    m->devs.assign(ctx->devs.size(), ModelDev());
    for (size_t i = 0; i < ctx->devs.size(); ++i) m->devs[i].dev = ctx->devs[i]->dev;
    ModelDev& md = m->devs[0];

    CU(cudaMalloc((void**)&md.xyz, 3 * N * sizeof(double)));
    CU(cudaMalloc((void**)&md.alpha, N * sizeof(double)));
    CU(cudaMalloc((void**)&m->label, N * sizeof(double)));
    CU(cudaMalloc((void**)&m->s2, N * sizeof(double)));
    CU(cudaMalloc((void**)&m->zfwd, N * sizeof(double)));
    CU(big_alloc(ctx, dc->dev, (void**)&m->L, N * N * sizeof(double)));
    CU(cudaMalloc((void**)&m->Dinv, (size_t)nb * TB * TB * sizeof(double)));
    CU(cudaMalloc((void**)&m->scratch, (8 + (size_t)nb * nb) * sizeof(int)));

    Workspace* ws = nullptr;
    int rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    struct Rel { DeviceCtx* d; Workspace* w; ~Rel() { ws_release(d, w); } } rel{dc, ws};
    cudaStream_t st = ws->st;

    // INT8-assisted factorisation (decided here so that its slice workspace is allocated before the timed phases start)
    bool want_i8 = false;
    int i8_panel = 0, i8_S = 7;
    if (!ctx->chol_serial) {
        static const long i8_min_n = getenv("GPR_FIT_INT8_MIN_N") ? atol(getenv("GPR_FIT_INT8_MIN_N")) : 8192;
        const char* fm = getenv("GPR_FIT_MODE");
        want_i8 = fm ? !strcmp(fm, "int8") : (long)n >= i8_min_n;
        i8_panel = getenv("GPR_FIT_PANEL") ? atoi(getenv("GPR_FIT_PANEL")) : std::max(8, std::min(32, nb / 8));
        if (const char* e = getenv("GPR_FIT_SLICES")) i8_S = std::max(6, std::min(8, atoi(e)));
        if (i8_panel < 1 || i8_panel >= nb) want_i8 = false;
    }
    struct I8Buf {                                               // returned to the context's buffer cache on every exit path
        gpr_ctx* c; int dev; signed char* p = nullptr; size_t bytes = 0; cudaStream_t s;
        ~I8Buf() { if (p) { cudaStreamSynchronize(s); big_free(c, dev, p, bytes); } }
    } i8_buf{ctx, dc->dev, nullptr, 0, st};
    if (want_i8) {
        i8_buf.bytes = ozaki_fit_workspace_bytes(i8_S, N);
        CU(big_alloc(ctx, dc->dev, (void**)&i8_buf.p, i8_buf.bytes));
        if (!ws->oz_ctrl) CU(cudaMalloc((void**)&ws->oz_ctrl, 4 * sizeof(int)));
    }

    // Internal point order.  Normally the caller's order (perm empty).  When the factorisation meets a
    // non-positive pivot with too many points after it for the trailing-block elimination, the offending point
    // is moved to the END of the internal order and the factorisation is retried: after at most MAX_TAIL moves
    // the offending points form the trailing block that gpr_tail.cu eliminates.  Everything on the device is in
    // internal order; alpha / normals are returned to the caller's order below.
    m->perm.clear();
    const char* no_tail_env = getenv("GPR_NO_TAIL");
    const bool tail_allowed = !(no_tail_env && atoi(no_tail_env) != 0);
    std::vector<double> gx, gy, gz, gl, gs;
    auto upload = [&]() -> int {
        const double *px = m->hx.data(), *py = m->hy.data(), *pz = m->hz.data(), *pl = m->hlabel.data();
        const double* ps = m->has_s2 ? m->hs2.data() : nullptr;
        if (!m->perm.empty()) {
            gx.resize(n); gy.resize(n); gz.resize(n); gl.resize(n); if (ps) gs.resize(n);
            for (size_t i = 0; i < n; ++i) {
                const size_t s = m->perm[i];
                gx[i] = px[s]; gy[i] = py[s]; gz[i] = pz[s]; gl[i] = pl[s]; if (ps) gs[i] = ps[s];
            }
            px = gx.data(); py = gy.data(); pz = gz.data(); pl = gl.data(); if (ps) ps = gs.data();
        }
        CU(cudaMemsetAsync(md.xyz, 0, 3 * N * sizeof(double), st));
        CU(cudaMemsetAsync(m->label, 0, N * sizeof(double), st));
        CU(cudaMemsetAsync(m->s2, 0, N * sizeof(double), st));
        // alpha of the padding points must be 0 (the thread-per-query kernel walks all N padded points); with an
        // indefinite tail the solves only write the rows of the leading block and of the tail
        CU(cudaMemsetAsync(md.alpha, 0, N * sizeof(double), st));
        CU(cudaMemcpyAsync(md.xyz, px, n * sizeof(double), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(md.xyz + N, py, n * sizeof(double), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(md.xyz + 2 * N, pz, n * sizeof(double), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m->label, pl, n * sizeof(double), cudaMemcpyHostToDevice, st));
        if (ps) CU(cudaMemcpyAsync(m->s2, ps, n * sizeof(double), cudaMemcpyHostToDevice, st));
        return GPR_OK;
    };
    CU(cudaEventRecord(ws->ev[0], st));
    rc = upload();
    if (rc) return rc;
    // the max-distance accumulator borrows the first 8 bytes of zfwd (reused before zfwd is written)
    unsigned long long* rbits = reinterpret_cast<unsigned long long*>(m->zfwd);
    CU(cudaMemsetAsync(rbits, 0, sizeof(unsigned long long), st));
    CU(cudaEventRecord(ws->ev[1], st));
    double Rbits_host = 0.0;
    int info[4] = {0, 0, 0, 0};
    size_t moved = 0;
    bool i8_failed = false;
    m->fit_int8_slices = 0;
    std::vector<int> counts;                 // conflict counts per point (caller's order), computed on the first failure
    for (;;) {
        CU(launch_cov_build(md.xyz, md.xyz + N, md.xyz + 2 * N, m->s2, (int)n, nb, 0, m->L, N, rbits, m->kp, st));
        if (moved == 0) {
            CU(cudaMemcpyAsync(&Rbits_host, rbits, sizeof(double), cudaMemcpyDeviceToHost, st));
            CU(cudaEventRecord(ws->ev[2], st));
        }
        // first attempt: every finished tile also goes to the registered peer replicas (gpr_ctx_set_fit_peers)
        const bool publish = moved == 0 && ctx->peers.n > 0 && ctx->peer_N == N && !ctx->chol_serial;
        // Large models: the flops left of each panel of tile columns run on the INT8 tensor cores (launch_cholesky_int8,
        // FP64-equivalent: 7 digit slices of base 254).  GPR_FIT_MODE=fp64|int8 forces a path, GPR_FIT_INT8_MIN_N (8192),
        // GPR_FIT_PANEL (tile columns per panel) and GPR_FIT_SLICES (7) tune it.  First attempt only: a matrix that turns out
        // not to be positive definite is re-factorised in FP64, whose pivot report the indefinite-tail logic below is tuned on.
        const bool use_i8 = want_i8 && moved == 0 && !i8_failed;
        if (use_i8) {
            signed char* Ls = i8_buf.p;
            cudaError_t ce = launch_cholesky_int8(m->L, N, nb, m->Dinv, m->scratch, dc->num_sms, st, publish ? &ctx->peers : nullptr, Ls,
                                                  i8_S, i8_panel, getenv("GPR_FIT_LAST") ? atoi(getenv("GPR_FIT_LAST")) : 48, ws->oz_ctrl);
            int octl[2] = {0, 0};
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(octl, ws->oz_ctrl, sizeof(octl), cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(info, m->scratch, sizeof(info), cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
            CU(ce);
            if (octl[1] != 0) return fail(GPR_ERR_CUDA, "INT8 update kernel of the factorisation aborted (barrier wait timed out)");
            if (info[1] != 0 || info[2] != 0) { i8_failed = true; info[1] = info[2] = 0; continue; }
            m->fit_int8_slices = i8_S;
        } else {
            CU(launch_cholesky(m->L, N, nb, m->Dinv, m->scratch, dc->num_sms, ctx->chol_serial, st, nullptr, publish ? &ctx->peers : nullptr));
            CU(cudaMemcpyAsync(info, m->scratch, sizeof(info), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        ctx->last_fit_published = publish && info[1] == 0;
        if (info[1] == 0) break;                                   // positive definite
        const size_t p = (size_t)info[1] - 1;                      // internal index of the failing pivot
        // Offending points are moved ("evicted") to the end of the internal order; once the failing pivot lies
        // inside the evicted region, that region is the trailing block that gpr_tail.cu eliminates.
        bool go_tail = tail_allowed && moved > 0 && p >= n - moved;
        if (!go_tail) {
            if (!tail_allowed || p == 0 || moved >= MAX_TAIL) break;   // give up: reported below
            if (m->perm.empty()) { m->perm.resize(n); for (size_t i = 0; i < n; ++i) m->perm[i] = i; }
            // A pair (a, b) with K_ab^2 > K_aa K_bb cannot sit in one positive definite block (its 2x2 minor is
            // indefinite).  counts[i] = number of such partners of point i in the whole set (device kernel): an
            // outlier — for the thin-plate kernel a point farther than R from most others — has many.
            if (counts.empty()) {
                counts.resize(n);
                int* dcounts = reinterpret_cast<int*>(m->scratch + 8);      // the flag area is free between factorisations
                if ((size_t)nb * nb < n) { CU(cudaFree(m->scratch)); CU(cudaMalloc((void**)&m->scratch, (8 + std::max((size_t)nb * nb, n)) * sizeof(int))); dcounts = m->scratch + 8; }
                CU(launch_conflict_counts(md.xyz, N, m->s2, (int)n, dcounts, m->kp, st));
                CU(cudaMemcpyAsync(counts.data(), dcounts, n * sizeof(int), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));                                // counts are in INTERNAL order == caller's order here
            }
            auto cnt_of = [&](size_t internal) { return counts[m->perm[internal]]; };
            auto kval = [&](size_t a, size_t b2) {
                const size_t ua = m->perm[a], ub = m->perm[b2];
                const double dx = m->hx[ua] - m->hx[ub], dy = m->hy[ua] - m->hy[ub], dz = m->hz[ua] - m->hz[ub];
                const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
                double k = m->kp.kind == 0 ? 2 * d * d * d - 3 * m->kp.p0 * d * d + m->kp.R3 : m->kp.amp * std::exp(-d * m->kp.inv);
                if (a == b2 && m->has_s2) k += m->hs2[ua];
                return k;
            };
            // the earlier point q that conflicts most with the failing pivot p; evict the one of the two with more conflicts
            size_t evict = p, q = p;
            double worst = 1.0;
            const double kpp = kval(p, p);
            for (size_t j = 0; j < p; ++j) {
                const double k = kval(p, j), r = k * k / (kpp * kval(j, j));
                if (r > worst) { worst = r; q = j; }
            }
            if (q != p && cnt_of(q) > cnt_of(p)) evict = q;
            if (cnt_of(evict) == 0) {
                // No point is to blame (no indefinite 2x2 minor): the matrix is indefinite as a whole — e.g. a nearly
                // noise-free thin-plate matrix, which is only conditionally positive definite.  Moving single points
                // away would let the factorisation run on into tiny pivots; take everything from the failing pivot on
                // as the trailing block if that is small enough, otherwise give up.
                if (n - p <= MAX_TAIL) go_tail = true; else break;
            }
          if (!go_tail) {
            // evict it together with every point that is at least half as conflicted (the other outliers), most
            // conflicted first, keeping the relative order of the rest
            const int thr = std::max(1, cnt_of(evict) / 2);
            std::vector<size_t> out;
            for (size_t i = 0; i < n - moved; ++i) if (i == evict || counts[m->perm[i]] >= thr) out.push_back(i);
            std::stable_sort(out.begin(), out.end(), [&](size_t a2, size_t b2) { return cnt_of(a2) > cnt_of(b2); });
            if (out.size() > MAX_TAIL - moved) {
                out.resize(MAX_TAIL - moved);
                if (std::find(out.begin(), out.end(), evict) == out.end()) out.back() = evict;
            }
            std::vector<char> gone(n, 0);
            for (size_t i : out) gone[i] = 1;
            std::vector<size_t> np;
            np.reserve(n);
            for (size_t i = 0; i < n - moved; ++i) if (!gone[i]) np.push_back(m->perm[i]);
            const size_t first_evicted = np.size();
            std::sort(out.begin(), out.end());
            for (size_t i : out) np.push_back(m->perm[i]);
            for (size_t i = n - moved; i < n; ++i) np.push_back(m->perm[i]);
            const bool unchanged = np == m->perm;
            m->perm.swap(np);
            moved += out.size();
            // the evicted points were already a suffix and the failure was at its first point: no retry needed
            go_tail = unchanged && p == first_evicted;
            if (!go_tail) {
                rc = upload();
                if (rc) return rc;
                continue;
            }
          }
        }
        {
            // Indefinite matrix whose offending points are now the last few of the internal order (in the node's own
            // ordering they are from the start: its external sphere points): keep the Cholesky factor of
            // the leading p points, eliminate the rest as one dense pivot block (gpr_tail.cu).  The leading block
            // is rebuilt with the tail rows as identity padding.
            m->n_spd = p; m->n_tail = n - p; m->nb = (int)((p + TB - 1) / TB);
            CU(launch_cov_build(md.xyz, md.xyz + N, md.xyz + 2 * N, m->s2, (int)p, m->nb, 0, m->L, N, rbits + 1, m->kp, st));
            CU(launch_cholesky(m->L, N, m->nb, m->Dinv, m->scratch, dc->num_sms, ctx->chol_serial, st));
            CU(cudaMemcpyAsync(info, m->scratch, sizeof(info), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            break;
        }
    }
    if (info[1] != 0) {
        const size_t user_pivot = m->perm.empty() ? (size_t)info[1] : m->perm[(size_t)info[1] - 1] + 1;
        g_pivot = (long long)user_pivot;
        char b[256];
        snprintf(b, sizeof b, "covariance matrix is not positive definite: pivot %zu of %zu is <= 0 "
                 "(thin-plate R must be >= the largest pairwise distance, %.6g here)", user_pivot, n, Rbits_host);
        m->perm.clear();
        return fail(GPR_ERR_NOT_SPD, b);
    }
    if (info[2] != 0) return fail(GPR_ERR_CUDA, "cholesky kernel aborted (dependency wait timed out)");
    if (!keep_R) m->R = Rbits_host;
    CU(cudaEventRecord(ws->ev[3], st));
    if (m->n_tail == 0) {
        CU(launch_trsv(0, m->L, N, nb, m->Dinv, m->label, m->zfwd, m->scratch, dc->num_sms, st));
        CU(launch_trsv(1, m->L, N, nb, m->Dinv, m->zfwd, md.alpha, m->scratch, dc->num_sms, st));
        // iterative refinement with a double-double residual (gpr_solve.cu); GPR_REFINE=0 switches it off
        for (int it = 0; it < ctx->refine_steps; ++it) {
            rc = ws_reserve(&ws->mpart, &ws->mpart_dbl, residual_scratch_doubles((int)N) + 2 * N);
            if (rc) return rc;
            double* rr = ws->mpart + residual_scratch_doubles((int)N);
            double* dd = rr + N;
            CU(launch_residual(md.xyz, N, m->s2, m->label, md.alpha, (int)n, (int)N, ws->mpart, rr, m->kp, st));
            CU(launch_trsv(0, m->L, N, nb, m->Dinv, rr, m->zfwd, m->scratch, dc->num_sms, st));
            CU(launch_trsv(1, m->L, N, nb, m->Dinv, m->zfwd, dd, m->scratch, dc->num_sms, st));
            CU(launch_axpy1(md.alpha, dd, (int)n, st));
        }
    } else {
        rc = tail_eliminate(m, dc, ws, true);
        if (rc) return rc;
    }
    CU(cudaEventRecord(ws->ev[4], st));
    m->h_alpha.resize(n);
    CU(cudaMemcpyAsync(m->h_alpha.data(), md.alpha, n * sizeof(double), cudaMemcpyDeviceToHost, st));   // internal order; un-permuted below
    m->h_normals.clear();
    m->n_normals = 0;
    if (m->with_normals) {
        // create<true>(): N_i = normalize(sum_j alpha_j k~(D_ij)(p_i - p_j))  (gp_regressor.hpp:166-181)
        rc = ws_reserve(&ws->io, &ws->io_cap, 14 * N);
        if (rc) return rc;
        double* f = ws->io; double* g = ws->io + N;
        const int nsplit = n <= 4096 ? 1 : predict_split((int)n, (int)N, dc->num_sms);
        if (nsplit > 1) { rc = ws_reserve(&ws->mpart, &ws->mpart_dbl, predict_part_doubles((int)n, (int)N)); if (rc) return rc; }
        CU(launch_predict(md.xyz, md.xyz + N, md.xyz + 2 * N, md.alpha, (int)n, (int)N, md.xyz, md.xyz + N, md.xyz + 2 * N,
                          (int)n, f, g, N, nullptr, 0, (int)n, m->kp, n <= 4096, ws->mpart, nsplit, st));
        CU(launch_normalize_rows(g, N, (int)n, st));
        m->h_normals.resize(3 * n);
        m->n_normals = n;
        for (int c = 0; c < 3; ++c)
            CU(cudaMemcpyAsync(m->h_normals.data() + c * n, g + c * N, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaEventRecord(ws->ev[5], st));
    CU(cudaStreamSynchronize(st));
    if (!m->perm.empty()) {                      // back to the caller's point order
        std::vector<double> tmp(m->h_alpha);
        for (size_t i = 0; i < n; ++i) m->h_alpha[m->perm[i]] = tmp[i];
        if (!m->h_normals.empty()) {
            tmp = m->h_normals;
            for (int c = 0; c < 3; ++c)
                for (size_t i = 0; i < n; ++i) m->h_normals[c * n + m->perm[i]] = tmp[c * n + i];
        }
    }
    int abortflag[4] = {0, 0, 0, 0};
    CU(cudaMemcpy(abortflag, m->scratch, sizeof(abortflag), cudaMemcpyDeviceToHost));
    if (abortflag[2] != 0) return fail(GPR_ERR_CUDA, "triangular solve kernel aborted (dependency wait timed out)");
    md.have = true;
    md.lfac = m->L; md.dinv = m->Dinv; md.own_fac = false; md.have_fac = m->n_tail == 0;
    {
        std::lock_guard<std::mutex> lk(ctx->tmu);
        gpr_timings& t = ctx->timings;
        t.h2d_ms = ev_ms(ws->ev[0], ws->ev[1]);
        t.cov_ms = ev_ms(ws->ev[1], ws->ev[2]);
        t.chol_ms = ev_ms(ws->ev[2], ws->ev[3]);
        t.solve_ms = ev_ms(ws->ev[3], ws->ev[4]);
        t.normals_ms = ev_ms(ws->ev[4], ws->ev[5]);
        t.fit_total_ms = ev_ms(ws->ev[1], ws->ev[5]);
        t.fit_int8_slices = (double)m->fit_int8_slices;
    }
    return GPR_OK;
}

// Build L^-1 on the primary device (lazily, for the variance path).  Caller holds m->mu.
static int ensure_linv_primary(gpr_model* m) {
    ModelDev& md = m->devs[0];
    if (md.have_linv) return GPR_OK;
    if (m->replica && !(md.have_fac && md.lfac && md.dinv)) return fail(GPR_ERR_INVALID, "replica model was created without L^-1 and without the factor");
    DeviceCtx* dc = m->ctx->devs[0];
    CU(cudaSetDevice(dc->dev));
    if (!md.linv) CU(big_alloc(m->ctx, dc->dev, (void**)&md.linv, m->cap * m->cap * sizeof(double)));
    // a replica that received the factor builds its own inverse from it (every GPU in parallel, nothing to exchange)
    const double* Lsrc = m->L ? m->L : md.lfac;
    const double* Dsrc = m->Dinv ? m->Dinv : md.dinv;
    if (!m->scratch) CU(cudaMalloc((void**)&m->scratch, (8 + (size_t)(m->cap / TB) * (m->cap / TB)) * sizeof(int)));
    Workspace* ws = nullptr;
    int rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    struct Rel { DeviceCtx* d; Workspace* w; ~Rel() { ws_release(d, w); } } rel{dc, ws};
    CU(cudaEventRecord(ws->ev[0], ws->st));
    CU(launch_linv(Lsrc, md.linv, m->cap, m->nb, Dsrc, m->scratch, dc->num_sms, ws->st));
    if (m->cap > m->N) CU(launch_identity_rows(md.linv, nullptr, m->cap, (int)m->N, (int)m->cap, (int)m->cap, ws->st));
    CU(cudaEventRecord(ws->ev[1], ws->st));
    int flags[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(flags, m->scratch, sizeof(flags), cudaMemcpyDeviceToHost, ws->st));
    CU(cudaStreamSynchronize(ws->st));
    if (flags[2] != 0) return fail(GPR_ERR_CUDA, "L^-1 kernel aborted (dependency wait timed out)");
    md.have_linv = true;
    std::lock_guard<std::mutex> lk(m->ctx->tmu);
    m->ctx->timings.linv_ms = ev_ms(ws->ev[0], ws->ev[1]);
    return GPR_OK;
}

// Make sure device slot di holds the predict state (copied from the primary over NVLink).
static int ensure_on_device(gpr_model* m, size_t di, bool need_linv, bool need_fac = false) {
    std::lock_guard<std::mutex> lk(m->mu);
    if (need_linv) { int rc = ensure_linv_primary(m); if (rc) return rc; }
    if (need_fac && !m->devs[0].have_fac) return fail(GPR_ERR_INVALID, "model holds no Cholesky factor on its primary device");
    if (di == 0) return GPR_OK;
    ModelDev& src = m->devs[0];
    ModelDev& dst = m->devs[di];
    CU(cudaSetDevice(dst.dev));
    if (!dst.have) {
        if (!dst.xyz) CU(cudaMalloc((void**)&dst.xyz, 3 * m->cap * sizeof(double)));
        if (!dst.alpha) CU(cudaMalloc((void**)&dst.alpha, m->cap * sizeof(double)));
        CU(cudaMemcpyPeer(dst.xyz, dst.dev, src.xyz, src.dev, 3 * m->cap * sizeof(double)));
        CU(cudaMemcpyPeer(dst.alpha, dst.dev, src.alpha, src.dev, m->cap * sizeof(double)));
        dst.have = true;
    }
    if (need_linv && !dst.have_linv) {
        if (!dst.linv) CU(big_alloc(m->ctx, dst.dev, (void**)&dst.linv, m->cap * m->cap * sizeof(double)));
        CU(cudaMemcpyPeer(dst.linv, dst.dev, src.linv, src.dev, m->cap * m->cap * sizeof(double)));
        dst.have_linv = true;
    }
    if (need_fac && !dst.have_fac) {
        const size_t db = (size_t)m->nb * TB * TB * sizeof(double);
        if (!dst.lfac) CU(big_alloc(m->ctx, dst.dev, (void**)&dst.lfac, m->cap * m->cap * sizeof(double)));
        if (!dst.dinv) CU(cudaMalloc((void**)&dst.dinv, (size_t)(m->cap / TB) * TB * TB * sizeof(double)));
        dst.own_fac = true;
        CU(cudaMemcpyPeer(dst.lfac, dst.dev, src.lfac, src.dev, m->cap * m->cap * sizeof(double)));
        CU(cudaMemcpyPeer(dst.dinv, dst.dev, src.dinv, src.dev, db));
        dst.have_fac = true;
    }
    if (need_linv && m->n_tail > 0 && !dst.have_tail) {
        const size_t zb = (size_t)(m->mp / 32) * m->cap * 32 * sizeof(double), sb = (size_t)m->mp * m->mp * sizeof(double);
        if (!dst.tZ) CU(cudaMalloc((void**)&dst.tZ, zb));
        if (!dst.tSinv) CU(cudaMalloc((void**)&dst.tSinv, sb));
        CU(cudaMemcpyPeer(dst.tZ, dst.dev, src.tZ, src.dev, zb));
        CU(cudaMemcpyPeer(dst.tSinv, dst.dev, src.tSinv, src.dev, sb));
        dst.have_tail = true;
    }
    return GPR_OK;
}

// Slices of X = L^-1 for the INT8 tensor-core variance (gpr_ozaki.cu), built once per model, device and slice count.
static int ensure_ozaki_slices(gpr_model* m, size_t di, int S, int base254, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(m->mu);
    ModelDev& md = m->devs[di];
    if (md.oz_xs && md.oz_S == S && md.oz_base == base254 && md.oz_ld == m->cap && md.oz_n == m->n) return GPR_OK;     // an append changes n: re-slice
    CU(cudaSetDevice(md.dev));
    // Same model, other slice count (escalation after a failed spot check): predict calls of other threads may be running on
    // the old buffers, so they are retired (freed with the model) instead of freed now.  A changed model (append / growth:
    // exclusive by the API's contract) frees them at once.
    const bool same_model = md.oz_n == m->n && md.oz_ld == m->cap;
    for (void* p : {(void*)md.oz_xs, (void*)md.oz_scale, (void*)md.oz_nz}) {
        if (!p) continue;
        if (same_model) m->retired.emplace_back(md.dev, p); else cudaFree(p);
    }
    md.oz_xs = nullptr; md.oz_scale = nullptr; md.oz_nz = nullptr; md.oz_S = 0;
    const size_t ld = m->cap;
    CU(cudaMalloc((void**)&md.oz_xs, (size_t)S * ld * ld));
    CU(cudaMalloc((void**)&md.oz_scale, 2 * ld * sizeof(double)));          // scales | row-max scratch
    CU(launch_ozaki_slice_x(md.linv, ld, m->nb * TB, S, base254, md.oz_xs, md.oz_scale,
                            reinterpret_cast<unsigned long long*>(md.oz_scale + ld), st));
    // which digit slices vanish entirely in which (row tile, k-block): their MMAs are skipped (gpr_ozaki.cu)
    const int nrt = m->nb, kbl = (int)(ld / 64);
    CU(cudaMalloc((void**)&md.oz_nz, (size_t)nrt * kbl));
    CU(launch_ozaki_mask(md.oz_xs, ld, ld * ld, S, nrt, kbl, md.oz_nz, (size_t)kbl, st));
    std::vector<unsigned char> hnz((size_t)nrt * kbl);
    CU(cudaMemcpyAsync(hnz.data(), md.oz_nz, hnz.size(), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    {
        double issued = 0.0, total = 0.0;
        for (int rt = 0; rt < nrt; ++rt)
            for (int kb = 0; kb < 2 * (rt + 1); ++kb)
                for (int t = 0; t < S; ++t) {
                    const double pairs = (double)(S - t);                       // pairs (t, u) with t + u < S
                    total += pairs;
                    if (kb == 0 || ((hnz[(size_t)rt * kbl + kb] >> t) & 1)) issued += pairs;
                }
        md.oz_exec = total > 0 ? issued / total : 1.0;
    }
    md.oz_S = S; md.oz_base = base254; md.oz_ld = ld; md.oz_n = m->n;
    return GPR_OK;
}

// ------------------------------------------------------------------------------------------------
// predict on one device.  in_dev/out_dev: pointers are device memory of that device.
// ------------------------------------------------------------------------------------------------
struct PredictIO {
    const double* qx; const double* qy; const double* qz; size_t q;
    double* f; double* var; double* grad; double* tx; double* ty;
    size_t out_ld;          // leading dimension of grad/tx/ty in the caller's arrays
    size_t offset;          // first query handled here
    bool device_ptrs;
    size_t q_call = 0;      // queries of the whole user call (the variance form is chosen on it, so that sharding a call over
                            // the devices of a context never changes the form); 0 = q
};

static int predict_on_device(gpr_model* m, size_t di, const PredictIO& io, double* mean_ms, double* var_ms,
                             double* h2d_ms, double* d2h_ms) {
    gpr_ctx* ctx = m->ctx;
    DeviceCtx* dc = ctx->devs[di];
    const bool want_var = io.var != nullptr, want_grad = io.grad != nullptr, want_t = io.tx != nullptr;
    // Which form of the variance (same n^2 flop per query on the FP64 tensor pipe either way):
    //   * forward substitution over L in the K* panel (var_trsm_kernel): needs no L^-1 — the default for large
    //     batches on a model whose inverse factor has not been built (time to first variance = the fit);
    //   * product with X = L^-1 (var_tiles_kernel / the fused q <= 8 kernel): no dependency chain, spreads a small
    //     batch over all SMs; used whenever X is resident anyway, for small batches and for indefinite-tail models.
    // GPR_VAR_MODE=trsm|product forces one (tests, bench).
    bool use_trsm = false, use_oz = false;
    int oz_S = 7, oz_base254 = 0;
    if (want_var && io.q > 8 && m->n_tail == 0) {
        // Three forms of the same n^2 flop per query:
        //   ozaki   : product with X = L^-1 on the INT8 tensor cores (tcgen05 kind::i8, gpr_ozaki.cu), FP64-equivalent by
        //             slicing (GPR_OZAKI_SLICES, default 7 slices of 7 bits: ~1e-9 of the variance; spot-checked against the
        //             FP64 product on every call) — ~2x the DMMA rate; needs X and its int8 slices (one-time per model);
        //   product : product with X on the FP64 tensor pipe (var_tiles_kernel);
        //   trsm    : forward substitution over L (var_trsm_kernel) — no X at all.
        // Default: ozaki for calls of >= GPR_OZAKI_MIN_Q (16384) queries (the one-time L^-1 + slicing pays off after about
        // one batch); otherwise product if X is resident, else trsm for >= GPR_TRSM_MIN_Q (4096) queries, else product.
        const char* mode_env = getenv("GPR_VAR_MODE");
        static const long oz_min_q = getenv("GPR_OZAKI_MIN_Q") ? atol(getenv("GPR_OZAKI_MIN_Q")) : 16384;
        static const long min_q = getenv("GPR_TRSM_MIN_Q") ? atol(getenv("GPR_TRSM_MIN_Q")) : 4096;
        // digit system: base 254 (|digit| <= 127, ~8 bits per slice: 6 slices).  Up to k = 22016 no int32 accumulator can
        // overflow; longer rows run the k-chunked kernel, which drains the accumulators into FP64 every 344 k-blocks.
        // GPR_OZAKI_BASE=128 selects |digit| <= 64 (7 slices, unchunked up to k = 74752); GPR_OZAKI_SLICES overrides the count.
        const long long kext = (long long)m->nb * TB;
        oz_base254 = 1;
        if (const char* e = getenv("GPR_OZAKI_BASE")) oz_base254 = atoi(e) == 254 ? 1 : 0;
        int bump;
        { std::lock_guard<std::mutex> lk(m->mu); bump = m->oz_bump; }
        oz_S = std::min(8, (oz_base254 ? 6 : 7) + bump);
        if (const char* e = getenv("GPR_OZAKI_SLICES")) oz_S = std::max(2, std::min(8, atoi(e) + bump));
        const bool oz_fits = ozaki_supported(oz_S, oz_base254, kext);
        const bool can_trsm = m->devs[0].have_fac;
        if (mode_env && !strcmp(mode_env, "ozaki")) use_oz = oz_fits;
        else if (mode_env && !strcmp(mode_env, "trsm")) use_trsm = can_trsm;
        else if (mode_env && !strcmp(mode_env, "product")) { }
        else {
            // sticky: once the int8 slices of a model exist every batch of more than 8 queries uses them, so that for a given
            // model state the result of a query does not depend on how the queries are split into calls
            bool have_x, dis, have_slices;
            { std::lock_guard<std::mutex> lk(m->mu); have_x = m->devs[0].have_linv; dis = m->oz_disabled; have_slices = m->devs[0].oz_xs != nullptr; }
            const long qc = (long)(io.q_call ? io.q_call : io.q);
            if (!dis && oz_fits && (qc >= oz_min_q || have_slices)) use_oz = true;
            else use_trsm = can_trsm && !have_x && qc >= min_q;
        }
    }
    int rc = ensure_on_device(m, di, want_var && !use_trsm, use_trsm);          // the INT8 path needs X = L^-1 too
    if (rc) return rc;
    CU(cudaSetDevice(dc->dev));
    ModelDev& md = m->devs[di];
    Workspace* ws = nullptr;
    rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    struct Rel { DeviceCtx* d; Workspace* w; ~Rel() { ws_release(d, w); } } rel{dc, ws};
    cudaStream_t st = ws->st;
    const size_t N = m->N, ld = m->cap;
    const int n = (int)m->n;
    if (!io.device_ptrs && io.q <= 8) {
        // The reference's callers: one query per call.  One fused launch, I/O through mapped pinned memory.
        if (!ws->hio) {
            CU(cudaHostAlloc((void**)&ws->hio, SMALL_HIO_DOUBLES * sizeof(double), cudaHostAllocMapped));
            CU(cudaHostGetDevicePointer((void**)&ws->hio_dev, ws->hio, 0));
        }
        const size_t Nv = (size_t)m->nb * TB;                   // padded rows of L^-1 (the leading block if there is a tail)
        const size_t need = predict_small_scratch_doubles((int)Nv);
        if (ws->small_dbl < need || ws->small_N != Nv) {
            rc = ws_reserve(&ws->small, &ws->small_dbl, need);
            if (rc) return rc;
            CU(cudaMemsetAsync(ws->small, 0, ws->small_dbl * sizeof(double), st));   // the layout (tickets) depends on Nv
            ws->small_N = Nv;
        }
        const int q = (int)io.q;
        for (int i = 0; i < q; ++i) {
            ws->hio[i] = io.qx[io.offset + i]; ws->hio[8 + i] = io.qy[io.offset + i]; ws->hio[16 + i] = io.qz[io.offset + i];
        }
        CU(cudaEventRecord(ws->ev[0], st));
        CU(launch_predict_small(md.xyz, md.xyz + ld, md.xyz + 2 * ld, md.alpha, n, (int)N, want_var ? md.linv : nullptr, ld,
                                ws->hio_dev, ws->small, q, want_var, want_grad, want_t, m->k0, m->kp, (int)m->n_spd, (int)Nv,
                                m->mp, md.tZ, md.tSinv, st));
        CU(cudaEventRecord(ws->ev[1], st));
        CU(cudaStreamSynchronize(st));
        const double* out = ws->hio + 24;
        for (int i = 0; i < q; ++i) {
            const size_t gi = io.offset + i;
            io.f[gi] = out[i];
            if (want_var) io.var[gi] = out[8 + i];
            for (int c = 0; c < 3; ++c) {
                if (want_grad) io.grad[c * io.out_ld + gi] = out[16 + 8 * c + i];
                if (want_t) { io.tx[c * io.out_ld + gi] = out[40 + 8 * c + i]; io.ty[c * io.out_ld + gi] = out[64 + 8 * c + i]; }
            }
        }
        const float t = ev_ms(ws->ev[0], ws->ev[1]);
        if (mean_ms) *mean_ms = want_var ? 0.0 : t;
        if (var_ms) *var_ms = want_var ? t : 0.0;
        if (h2d_ms) *h2d_ms = 0.0;
        if (d2h_ms) *d2h_ms = 0.0;
        return GPR_OK;
    }
    // workspace and model-side slices of the INT8 form for the current (oz_S, oz_base254)
    size_t oz_panel_ld_max = 0;
    auto prep_oz = [&]() -> int {
        const size_t need = (size_t)oz_S * oz_panel_ld_max * m->cap;
        if (ws->oz_ks_bytes < need) {
            if (ws->oz_ks) CU(cudaFree(ws->oz_ks));
            ws->oz_ks = nullptr; ws->oz_ks_bytes = 0;
            CU(cudaMalloc((void**)&ws->oz_ks, need));
            CU(cudaMemsetAsync(ws->oz_ks, 0, need, st));
            ws->oz_ks_bytes = need;
        }
        if (!ws->oz_ctrl) CU(cudaMalloc((void**)&ws->oz_ctrl, 4 * sizeof(int)));
        return ensure_ozaki_slices(m, di, oz_S, oz_base254, st);
    };
    const bool small_var = want_var && io.q <= 8;
    size_t batch = io.q;
    if (want_var && !small_var) batch = std::min(io.q, ctx->query_tile ? ctx->query_tile : (size_t)TB * dc->num_sms);
    else batch = std::min(io.q, (size_t)1 << 22);
    const size_t panel_ld_max = (batch + TB - 1) / TB * TB;
    oz_panel_ld_max = panel_ld_max;
    rc = ws_reserve(&ws->io, &ws->io_cap, 14 * std::max(batch, (size_t)TB));
    if (rc) return rc;
    const size_t cap = ws->io_cap / 14;
    double* dq = ws->io; double* df = dq + 3 * cap; double* dv = df + cap; double* dg = dv + cap;
    double* dtx = dg + 3 * cap; double* dty = dtx + 3 * cap;
    if (want_var) {
        rc = ws_reserve(&ws->panel, &ws->panel_dbl, N * panel_ld_max);
        if (rc) return rc;
        rc = ws_reserve(&ws->partial, &ws->partial_dbl, std::max((size_t)m->nb * panel_ld_max, (size_t)8 * 8 * N));
        if (rc) return rc;
        if (use_oz) {
            rc = prep_oz();
            if (rc) return rc;
        }
    }
    float t_mean = 0, t_var = 0, t_h2d = 0, t_d2h = 0, t_oz = 0;
    for (size_t b0 = 0; b0 < io.q; b0 += batch) {
        bool oz_timed = false;
        const size_t bq = std::min(batch, io.q - b0);
        const size_t g0 = io.offset + b0;
        const double *qx, *qy, *qz;
        double *f, *v, *g;
        size_t gld;
        CU(cudaEventRecord(ws->ev[0], st));
        if (io.device_ptrs) {
            qx = io.qx + g0; qy = io.qy + g0; qz = io.qz + g0;
            f = io.f + g0; v = want_var ? io.var + g0 : nullptr; g = want_grad ? io.grad + g0 : nullptr; gld = io.out_ld;
        } else {
            CU(cudaMemcpyAsync(dq, io.qx + g0, bq * sizeof(double), cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(dq + cap, io.qy + g0, bq * sizeof(double), cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(dq + 2 * cap, io.qz + g0, bq * sizeof(double), cudaMemcpyHostToDevice, st));
            qx = dq; qy = dq + cap; qz = dq + 2 * cap;
            f = df; v = want_var ? dv : nullptr; g = want_grad ? dg : nullptr; gld = cap;
        }
        CU(cudaEventRecord(ws->ev[1], st));
        const size_t pld = small_var ? (size_t)TB : (bq + TB - 1) / TB * TB;
        const bool warp_mode = small_var || (!want_var && bq <= (size_t)64 * dc->num_sms);
        int nsplit = 1;
        if (!warp_mode) {
            nsplit = predict_split((int)(want_var ? pld : bq), (int)N, dc->num_sms);
            if (nsplit > 1) { rc = ws_reserve(&ws->mpart, &ws->mpart_dbl, predict_part_doubles((int)bq, (int)N)); if (rc) return rc; }
        }
        CU(launch_predict(md.xyz, md.xyz + ld, md.xyz + 2 * ld, md.alpha, n, (int)N, qx, qy, qz, (int)bq, f, g, gld,
                          want_var ? ws->panel : nullptr, pld, (int)m->n_spd, m->kp, warp_mode, ws->mpart, nsplit, st));
        CU(cudaEventRecord(ws->ev[2], st));
        if (want_var) {
            if (small_var) CU(launch_variance_small(md.linv, ld, m->nb * TB, ws->panel, pld, (int)bq, ws->partial, m->k0, v, st));
            else if (use_oz) {
                // slice the K* panel (one power-of-two scale: |k*| <= k(0) for these kernels), multiply on the INT8 tensor
                // cores, recombine + column norms in the kernel's FP64 epilogue, then the usual fixed-order finalize
                int ge = 0;
                frexp(m->k0, &ge);
                const double cs = ldexp(1.0, ge);
                for (;;) {
                    CU(launch_ozaki_slice_panel(ws->panel, pld, (int)bq, m->nb * TB, 1.0 / cs, oz_S, oz_base254, ws->oz_ks, ld, pld, st));
                    CU(cudaEventRecord(ws->ev[5], st));
                    CU(launch_ozaki_product(md.oz_xs, ld, ld * ld, m->nb, ws->oz_ks, ld, pld * ld, pld, (int)bq, pld, (size_t)m->nb * TB, 1,
                                            oz_S, oz_base254, md.oz_scale, cs, ws->partial, ws->oz_ctrl, nullptr, 0, st, md.oz_nz, ld / 64));
                    CU(cudaEventRecord(ws->ev[6], st));
                    oz_timed = true;
                    CU(launch_var_finalize(ws->partial, pld, m->nb, (int)bq, m->k0, v, st));
                    if (b0 != 0) break;
                    // spot check of this call: the first query tile again on the FP64 tensor pipe (one tile: ~2 % of a batch)
                    const int cq = (int)std::min<size_t>(bq, TB);
                    rc = ws_reserve(&ws->tailw, &ws->tailw_dbl, (size_t)m->nb * TB + 2 * TB);
                    if (rc) return rc;
                    double* chk_part = ws->tailw; double* chk_v = chk_part + (size_t)m->nb * TB;
                    CU(launch_variance(md.linv, ld, m->nb, ws->panel, TB, cq, chk_part, m->k0, chk_v, st, pld));
                    double hv[2 * TB];
                    CU(cudaMemcpyAsync(hv, v, cq * sizeof(double), cudaMemcpyDeviceToHost, st));
                    CU(cudaMemcpyAsync(hv + TB, chk_v, cq * sizeof(double), cudaMemcpyDeviceToHost, st));
                    CU(cudaStreamSynchronize(st));
                    double dmax = 0.0, vmax = 0.0;
                    for (int i = 0; i < cq; ++i) { dmax = std::max(dmax, std::fabs(hv[i] - hv[TB + i])); vmax = std::max(vmax, std::fabs(hv[TB + i])); }
                    if (dmax <= 1e-8 * std::max(vmax, 1e-300)) break;
                    // Not accurate enough for this model (conditioning beyond what the slice count covers): one more slice
                    // (x128 / x254 finer) if the accumulators still cannot overflow — remembered on the model — else this batch
                    // and everything after it on this model go through the FP64 product form.
                    const bool forced_slices = getenv("GPR_OZAKI_SLICES") != nullptr;
                    const int nS = oz_S + 1;
                    if (!forced_slices && ozaki_supported(nS, oz_base254, (long long)m->nb * TB)) {
                        { std::lock_guard<std::mutex> lk(m->mu); m->oz_bump += 1; }
                        oz_S = nS;
                        rc = prep_oz();
                        if (rc) return rc;
                        continue;
                    }
                    { std::lock_guard<std::mutex> lk(m->mu); m->oz_disabled = true; }
                    use_oz = false;
                    CU(launch_variance(md.linv, ld, m->nb, ws->panel, pld, (int)bq, ws->partial, m->k0, v, st));
                    break;
                }
            }
            else if (use_trsm) CU(launch_variance_trsm(md.lfac, ld, m->nb, md.dinv, ws->panel, pld, (int)bq, ws->partial, m->k0, v, st));
            else CU(launch_variance(md.linv, ld, m->nb, ws->panel, pld, (int)bq, ws->partial, m->k0, v, st));
            if (m->n_tail > 0) {
                // indefinite tail: var -= w^T S^-1 w,  w = k_2 - Z^T k_1  (gpr_tail.cu)
                const int nslab = m->mp / 32;
                rc = ws_reserve(&ws->tailw, &ws->tailw_dbl, (size_t)nslab * pld * 32);
                if (rc) return rc;
                for (int s = 0; s < nslab; ++s)
                    CU(launch_skinny(2, ws->panel, pld, (int)bq, (int)m->n_spd, md.tZ + (size_t)s * ld * 32,
                                     ws->tailw + (size_t)s * pld * 32, st));
                CU(launch_tail_var(qx, qy, qz, (int)bq, md.xyz, ld, (int)m->n_spd, (int)m->n_tail, m->mp, ws->tailw, pld,
                                   md.tSinv, v, m->kp, st));
            }
        }
        CU(cudaEventRecord(ws->ev[3], st));
        if (want_t && !io.device_ptrs) CU(launch_tangent_basis(g, gld, (int)bq, dtx, dty, st));
        if (!io.device_ptrs) {
            CU(cudaMemcpyAsync(io.f + g0, f, bq * sizeof(double), cudaMemcpyDeviceToHost, st));
            if (want_var) CU(cudaMemcpyAsync(io.var + g0, v, bq * sizeof(double), cudaMemcpyDeviceToHost, st));
            for (int c = 0; c < 3; ++c) {
                if (want_grad) CU(cudaMemcpyAsync(io.grad + c * io.out_ld + g0, g + c * gld, bq * sizeof(double), cudaMemcpyDeviceToHost, st));
                if (want_t) {
                    CU(cudaMemcpyAsync(io.tx + c * io.out_ld + g0, dtx + c * gld, bq * sizeof(double), cudaMemcpyDeviceToHost, st));
                    CU(cudaMemcpyAsync(io.ty + c * io.out_ld + g0, dty + c * gld, bq * sizeof(double), cudaMemcpyDeviceToHost, st));
                }
            }
        }
        CU(cudaEventRecord(ws->ev[4], st));
        CU(cudaStreamSynchronize(st));
        t_h2d += ev_ms(ws->ev[0], ws->ev[1]); t_mean += ev_ms(ws->ev[1], ws->ev[2]);
        t_var += ev_ms(ws->ev[2], ws->ev[3]); t_d2h += ev_ms(ws->ev[3], ws->ev[4]);
        if (oz_timed) t_oz += ev_ms(ws->ev[5], ws->ev[6]);
    }
    {
        std::lock_guard<std::mutex> lk(ctx->tmu);
        ctx->timings.ozaki_ms = t_oz; ctx->timings.ozaki_slices = t_oz > 0 ? (double)oz_S : 0.0;
        ctx->timings.ozaki_issued_fraction = t_oz > 0 ? md.oz_exec : 0.0;
    }
    if (use_oz) {
        int ctrl[2] = {0, 0};
        CU(cudaMemcpy(ctrl, ws->oz_ctrl, sizeof ctrl, cudaMemcpyDeviceToHost));
        if (ctrl[1] != 0) return fail(GPR_ERR_CUDA, "INT8 tensor-core variance kernel aborted (a pipeline wait timed out)");
    }
    if (mean_ms) *mean_ms = t_mean;
    if (var_ms) *var_ms = t_var;
    if (h2d_ms) *h2d_ms = t_h2d;
    if (d2h_ms) *d2h_ms = t_d2h;
    return GPR_OK;
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* gpr_last_error(void) { return g_err.c_str(); }
long long gpr_last_pivot(void) { return g_pivot; }

int gpr_ctx_create(const int* devices, int ndev, gpr_ctx** out) {
    if (!out) return fail(GPR_ERR_INVALID, "null output pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(GPR_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    gpr_ctx* ctx = new gpr_ctx();
    memset(&ctx->timings, 0, sizeof ctx->timings);
    std::vector<int> devs;
    if (!devices || ndev <= 0) devs.push_back(0);
    else devs.assign(devices, devices + ndev);
    for (int d : devs) {
        if (d < 0 || d >= count) { delete ctx; return fail(GPR_ERR_INVALID, "device index out of range"); }
        cudaDeviceProp p;
        e = cudaGetDeviceProperties(&p, d);
        if (e != cudaSuccess) { delete ctx; return fail(GPR_ERR_CUDA, cudaGetErrorString(e)); }
        if (p.major < 10) { delete ctx; return fail(GPR_ERR_CUDA, "device is not sm_100 or newer; this library is built for sm_100a only"); }
        DeviceCtx* dc = new DeviceCtx();
        dc->dev = d; dc->num_sms = p.multiProcessorCount;
        ctx->devs.push_back(dc);
    }
    for (size_t a = 0; a < devs.size(); ++a)
        for (size_t b = 0; b < devs.size(); ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[a], devs[b]);
            if (can) { cudaSetDevice(devs[a]); cudaDeviceEnablePeerAccess(devs[b], 0); cudaGetLastError(); }
        }
    if (const char* s = getenv("GPR_QUERY_TILE")) {
        long v = atol(s);
        if (v > 0) ctx->query_tile = (size_t)(v + TB - 1) / TB * TB;
    }
    if (const char* s = getenv("GPR_CHOL_SERIAL")) ctx->chol_serial = atoi(s);
    if (const char* s = getenv("GPR_REFINE")) ctx->refine_steps = std::max(0, std::min(4, atoi(s)));
    { std::lock_guard<std::mutex> lk(g_live_mu); g_live_ctx.push_back(ctx); }
    *out = ctx;
    return GPR_OK;
}

int gpr_ctx_destroy(gpr_ctx* ctx) {
    if (!ctx) return GPR_OK;
    {
        std::lock_guard<std::mutex> lk(g_live_mu);
        g_live_ctx.erase(std::remove(g_live_ctx.begin(), g_live_ctx.end(), ctx), g_live_ctx.end());
    }
    gpr_ctx_clear_fit_peers(ctx);
    for (auto& b : ctx->big_cache) { cudaSetDevice(b.dev); cudaFree(b.p); }
    ctx->big_cache.clear();
    for (DeviceCtx* dc : ctx->devs) {
        cudaSetDevice(dc->dev);
        for (Workspace* ws : dc->free_ws) {
            cudaFree(ws->io); cudaFree(ws->panel); cudaFree(ws->partial); cudaFree(ws->mpart); cudaFree(ws->small); cudaFree(ws->tailw);
            cudaFree(ws->oz_ks); cudaFree(ws->oz_ctrl);
            if (ws->hio) cudaFreeHost(ws->hio);
            for (auto& e : ws->ev) cudaEventDestroy(e);
            cudaStreamDestroy(ws->st);
            delete ws;
        }
        delete dc;
    }
    delete ctx;
    return GPR_OK;
}

int gpr_ctx_num_devices(const gpr_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int gpr_ctx_clear_fit_peers(gpr_ctx* ctx) {
    if (!ctx) return fail(GPR_ERR_INVALID, "null context");
    if (!ctx->peer_maps.empty()) {
        CU(cudaSetDevice(ctx->devs[0]->dev));
        for (void* p : ctx->peer_maps) cudaIpcCloseMemHandle(p);
    }
    ctx->peer_maps.clear();
    ctx->peers = CholPeers{};
    ctx->peer_N = 0;
    return GPR_OK;
}

int gpr_model_ipc_export(gpr_ctx* ctx, gpr_model* m, void* handles128) {
    if (!ctx || !m || !handles128) return fail(GPR_ERR_INVALID, "null pointer");
    if (!m->replica || !m->devs[0].lfac || !m->devs[0].dinv || !m->devs[0].own_fac)
        return fail(GPR_ERR_INVALID, "gpr_model_ipc_export needs a replica created with the factor buffers (with_linv & 2)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CU(cudaSetDevice(m->devs[0].dev));
    cudaIpcMemHandle_t h[2];
    CU(cudaIpcGetMemHandle(&h[0], m->devs[0].lfac));
    CU(cudaIpcGetMemHandle(&h[1], m->devs[0].dinv));
    memcpy(handles128, h, sizeof h);
    return GPR_OK;
}

int gpr_ctx_set_fit_peers(gpr_ctx* ctx, const void* handles128, int n_peers, size_t n) {
    if (!ctx || (n_peers > 0 && !handles128)) return fail(GPR_ERR_INVALID, "null pointer");
    if (n_peers < 0 || n_peers > MAX_CHOL_PEERS) return fail(GPR_ERR_INVALID, "at most 7 peers");
    int rc = gpr_ctx_clear_fit_peers(ctx);
    if (rc) return rc;
    if (n_peers == 0) return GPR_OK;
    CU(cudaSetDevice(ctx->devs[0]->dev));
    const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles128);
    for (int i = 0; i < n_peers; ++i) {
        void *pl = nullptr, *pd = nullptr;
        cudaIpcMemHandle_t hl, hd;
        memcpy(&hl, &h[2 * i], sizeof hl); memcpy(&hd, &h[2 * i + 1], sizeof hd);
        cudaError_t e = cudaIpcOpenMemHandle(&pl, hl, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) { ctx->peer_maps.push_back(pl); e = cudaIpcOpenMemHandle(&pd, hd, cudaIpcMemLazyEnablePeerAccess); }
        if (e != cudaSuccess) {
            cudaGetLastError();
            gpr_ctx_clear_fit_peers(ctx);
            return fail(GPR_ERR_CUDA, std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e));
        }
        ctx->peer_maps.push_back(pd);
        ctx->peers.L[i] = static_cast<double*>(pl); ctx->peers.Dinv[i] = static_cast<double*>(pd);
    }
    ctx->peers.n = n_peers;
    ctx->peer_N = (n + TB - 1) / TB * TB;
    return GPR_OK;
}

int gpr_ctx_last_fit_published(const gpr_ctx* ctx) { return ctx && ctx->last_fit_published ? 1 : 0; }

int gpr_last_timings(const gpr_ctx* ctx, gpr_timings* out) {
    if (!ctx || !out) return fail(GPR_ERR_INVALID, "null pointer");
    std::lock_guard<std::mutex> lk(const_cast<gpr_ctx*>(ctx)->tmu);
    *out = ctx->timings;
    return GPR_OK;
}

int gpr_fit(gpr_ctx* ctx, const double* x, const double* y, const double* z, const double* label,
            const double* sigma2, size_t n, gpr_kernel_t kernel, int with_normals, gpr_model** out) {
    if (!ctx || !out) return fail(GPR_ERR_INVALID, "null context or output pointer");
    if (!x || !y || !z || !label || n == 0) return fail(GPR_ERR_INVALID, "All input data is empty!");
    if (kernel.kind < 0 || kernel.kind > 2) return fail(GPR_ERR_INVALID, "unknown kernel kind");
    gpr_model* m = new gpr_model();
    m->ctx = ctx; m->kernel = kernel; m->kp = make_kp(kernel); m->k0 = kernel_at_zero(m->kp);
    m->has_s2 = sigma2 != nullptr; m->with_normals = with_normals != 0;
    m->hx.assign(x, x + n); m->hy.assign(y, y + n); m->hz.assign(z, z + n); m->hlabel.assign(label, label + n);
    if (sigma2) m->hs2.assign(sigma2, sigma2 + n);
    int rc = fit_from_host(m, false);
    if (rc) { free_factor(m); delete m; return rc; }
    *out = m;
    return GPR_OK;
}

int gpr_model_destroy(gpr_model* m) {
    if (!m) return GPR_OK;
    free_factor(m);
    delete m;
    return GPR_OK;
}

size_t gpr_model_size(const gpr_model* m) { return m ? m->n : 0; }
size_t gpr_model_tail_size(const gpr_model* m) { return m ? m->n_tail : 0; }

int gpr_model_get(const gpr_model* m, double* alpha, double* R, double* normals) {
    if (!m) return fail(GPR_ERR_INVALID, "Empty Model pointer");
    if (alpha) {
        if (m->h_alpha.size() != m->n) return fail(GPR_ERR_INVALID, "model holds no host alpha (replica)");
        memcpy(alpha, m->h_alpha.data(), m->n * sizeof(double));
    }
    if (R) *R = m->R;
    if (normals) {
        // Normals exist for the points of the last create<true>(); rows appended by update() stay zero
        // (the reference does not refresh them either, gp_regressor.hpp:462-477).
        if (m->n_normals == 0 || m->h_normals.size() != 3 * m->n_normals) return fail(GPR_ERR_INVALID, "model was fitted without normals");
        for (int c = 0; c < 3; ++c) {
            memcpy(normals + c * m->n, m->h_normals.data() + c * m->n_normals, m->n_normals * sizeof(double));
            for (size_t i = m->n_normals; i < m->n; ++i) normals[c * m->n + i] = 0.0;
        }
    }
    return GPR_OK;
}

int gpr_model_get_factor(const gpr_model* m, double* L) {
    if (!m || !L) return fail(GPR_ERR_INVALID, "null pointer");
    if (!m->L) return fail(GPR_ERR_INVALID, "model holds no factor (replica)");
    CU(cudaSetDevice(m->devs[0].dev));
    CU(cudaMemcpy2D(L, m->n * sizeof(double), m->L, m->cap * sizeof(double), m->n * sizeof(double), m->n, cudaMemcpyDeviceToHost));
    for (size_t c = 0; c < m->n; ++c)
        for (size_t r = 0; r < c; ++r) L[c * m->n + r] = 0.0;
    return GPR_OK;
}

}  // extern "C"

static int predict_host(gpr_ctx* ctx, gpr_model* m, const double* qx, const double* qy, const double* qz, size_t q,
                        double* f, double* var, double* grad, double* tx, double* ty);

// One combined launch for the small requests collected by the micro-batcher: their queries are gathered into one SoA
// batch, evaluated through the regular batched path (so every request gets exactly the bits a batched call of the
// same queries returns) and scattered back.  A batch of one request is passed through unchanged (fused q <= 8 kernel).
static void run_small_batch(gpr_ctx* ctx, gpr_model* m, std::vector<gpr_model::SmallReq*>& reqs) {
    if (reqs.size() == 1) {
        gpr_model::SmallReq* r = reqs[0];
        r->rc = predict_host(ctx, m, r->qx, r->qy, r->qz, r->q, r->f, r->var, r->grad, r->tx, r->ty);
        if (r->rc) r->err = g_err;
        return;
    }
    size_t total = 0;
    bool want_var = false, want_grad = false, want_t = false;
    for (auto* r : reqs) { total += r->q; want_var |= r->var != nullptr; want_grad |= r->grad != nullptr; want_t |= r->tx != nullptr; }
    std::vector<double> in(3 * total), out((1 + (want_var ? 1 : 0) + (want_grad ? 3 : 0) + (want_t ? 6 : 0)) * total);
    size_t o = 0;
    for (auto* r : reqs) {
        for (size_t i = 0; i < r->q; ++i) { in[o + i] = r->qx[i]; in[total + o + i] = r->qy[i]; in[2 * total + o + i] = r->qz[i]; }
        o += r->q;
    }
    double* bf = out.data();
    double* bv = want_var ? bf + total : nullptr;
    double* bg = want_grad ? bf + (1 + (want_var ? 1 : 0)) * total : nullptr;
    double* btx = want_t ? bg + 3 * total : nullptr;
    double* bty = want_t ? btx + 3 * total : nullptr;
    const int rc = predict_host(ctx, m, in.data(), in.data() + total, in.data() + 2 * total, total, bf, bv, bg, btx, bty);
    const std::string err = rc ? g_err : std::string();
    o = 0;
    for (auto* r : reqs) {
        r->rc = rc; r->err = err;
        if (!rc)
            for (size_t i = 0; i < r->q; ++i) {
                r->f[i] = bf[o + i];
                if (r->var) r->var[i] = bv[o + i];
                for (int c = 0; c < 3; ++c) {
                    if (r->grad) r->grad[c * r->out_ld + i] = bg[c * total + o + i];
                    if (r->tx) { r->tx[c * r->out_ld + i] = btx[c * total + o + i]; r->ty[c * r->out_ld + i] = bty[c * total + o + i]; }
                }
            }
        o += r->q;
    }
}

// Flat combining: the first thread to arrive becomes the leader and runs its own request at once (a lone caller sees
// the latency of the fused single-query kernel); requests that arrive while a launch is in flight queue up and are
// served together by the next launch (at most MICRO_CAP queries).  After MICRO_ROUNDS launches the leader hands over
// to a waiting thread, so that it cannot be kept working for others forever.
constexpr size_t MICRO_CAP = 1024;
constexpr int MICRO_ROUNDS = 4;
static bool microbatch_on() { static const bool on = !(getenv("GPR_MICROBATCH") && atoi(getenv("GPR_MICROBATCH")) == 0); return on; }

static int predict_small_combined(gpr_ctx* ctx, gpr_model* m, const double* qx, const double* qy, const double* qz, size_t q,
                                  double* f, double* var, double* grad, double* tx, double* ty) {
    gpr_model::SmallReq me;
    me.qx = qx; me.qy = qy; me.qz = qz; me.q = q; me.f = f; me.var = var; me.grad = grad; me.tx = tx; me.ty = ty; me.out_ld = q;
    std::unique_lock<std::mutex> lk(m->bmu);
    m->pending.push_back(&me);
    for (;;) {
        if (me.done) break;
        if (m->leader) { me.cv.wait(lk, [&] { return me.done || !m->leader; }); continue; }
        m->leader = true;
        for (int round = 0; round < MICRO_ROUNDS && !m->pending.empty(); ++round) {
            std::vector<gpr_model::SmallReq*> take;
            size_t total = 0, cnt = 0;
            while (cnt < m->pending.size() && total + m->pending[cnt]->q <= MICRO_CAP) total += m->pending[cnt++]->q;
            take.assign(m->pending.begin(), m->pending.begin() + (std::ptrdiff_t)cnt);
            m->pending.erase(m->pending.begin(), m->pending.begin() + (std::ptrdiff_t)cnt);
            lk.unlock();
            run_small_batch(ctx, m, take);
            lk.lock();
            for (auto* r : take) { r->done = true; if (r != &me) r->cv.notify_one(); }
        }
        m->leader = false;
        if (!m->pending.empty()) m->pending.front()->cv.notify_one();      // hand over: the oldest waiter becomes the leader
    }
    lk.unlock();
    if (me.rc) g_err = me.err;
    return me.rc;
}

extern "C" {

int gpr_predict(gpr_ctx* ctx, gpr_model* m, const double* qx, const double* qy, const double* qz, size_t q,
                double* f, double* var, double* grad, double* tx, double* ty) {
    if (!ctx) return fail(GPR_ERR_INVALID, "null context");
    if (!m) return fail(GPR_ERR_INVALID, "Empty Model pointer");
    if (!qx || !qy || !qz || !f || q == 0) return fail(GPR_ERR_INVALID, "All input data is empty!");
    if ((tx == nullptr) != (ty == nullptr) || (tx && !grad)) return fail(GPR_ERR_INVALID, "tx/ty need each other and grad");
    if (q <= 8 && microbatch_on()) return predict_small_combined(ctx, m, qx, qy, qz, q, f, var, grad, tx, ty);
    return predict_host(ctx, m, qx, qy, qz, q, f, var, grad, tx, ty);
}

}  // extern "C"

static int predict_host(gpr_ctx* ctx, gpr_model* m, const double* qx, const double* qy, const double* qz, size_t q,
                        double* f, double* var, double* grad, double* tx, double* ty) {
    const size_t nd = ctx->devs.size();
    // Shard contiguous query ranges over the devices; tiny batches stay on the primary device.
    const size_t use = (nd > 1 && q >= 4096 * nd) ? nd : 1;
    std::vector<int> rcs(use, 0);
    std::vector<std::string> errs(use);
    std::vector<double> tm(use, 0), tv(use, 0), th(use, 0), td(use, 0);
    auto work = [&](size_t di) {
        PredictIO io;
        const size_t a = q * di / use, b = q * (di + 1) / use;
        io.qx = qx; io.qy = qy; io.qz = qz; io.q = b - a; io.offset = a;
        io.f = f; io.var = var; io.grad = grad; io.tx = tx; io.ty = ty; io.out_ld = q; io.device_ptrs = false;
        io.q_call = q;
        rcs[di] = predict_on_device(m, di, io, &tm[di], &tv[di], &th[di], &td[di]);
        if (rcs[di]) errs[di] = g_err;
    };
    if (use == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (size_t di = 0; di < use; ++di) pool.emplace_back(work, di);
        for (auto& t : pool) t.join();
    }
    for (size_t di = 0; di < use; ++di) if (rcs[di]) return fail(rcs[di], errs[di]);
    std::lock_guard<std::mutex> lk(ctx->tmu);
    gpr_timings& t = ctx->timings;
    t.predict_mean_ms = *std::max_element(tm.begin(), tm.end());
    t.predict_var_ms = *std::max_element(tv.begin(), tv.end());
    t.h2d_ms = *std::max_element(th.begin(), th.end());
    t.d2h_ms = *std::max_element(td.begin(), td.end());
    t.predict_total_ms = t.predict_mean_ms + t.predict_var_ms + t.h2d_ms + t.d2h_ms;
    return GPR_OK;
}

extern "C" {

int gpr_predict_device(gpr_ctx* ctx, gpr_model* m, const double* d_qx, const double* d_qy, const double* d_qz,
                       size_t q, double* d_f, double* d_var, double* d_grad) {
    if (!ctx) return fail(GPR_ERR_INVALID, "null context");
    if (!m) return fail(GPR_ERR_INVALID, "Empty Model pointer");
    if (!d_qx || !d_qy || !d_qz || !d_f || q == 0) return fail(GPR_ERR_INVALID, "All input data is empty!");
    PredictIO io;
    io.qx = d_qx; io.qy = d_qy; io.qz = d_qz; io.q = q; io.offset = 0;
    io.f = d_f; io.var = d_var; io.grad = d_grad; io.tx = nullptr; io.ty = nullptr; io.out_ld = q; io.device_ptrs = true;
    double tm = 0, tv = 0, th = 0, td = 0;
    int rc = predict_on_device(m, 0, io, &tm, &tv, &th, &td);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->tmu);
    ctx->timings.predict_mean_ms = tm; ctx->timings.predict_var_ms = tv;
    ctx->timings.predict_total_ms = tm + tv;
    return GPR_OK;
}

// Lattice points [g_begin, g_end) of the sampler on device di: coordinates generated on the device, mean through K4,
// compaction of |f| <= tol on the device; only the survivors (global index, f) come back.
static int sample_range(gpr_model* m, size_t di, const std::vector<double>& axis, unsigned long long g_begin,
                        unsigned long long g_end, double tol, std::vector<std::pair<unsigned long long, double>>& hits,
                        double* ms) {
    gpr_ctx* ctx = m->ctx;
    int rc = ensure_on_device(m, di, false);
    if (rc) return rc;
    DeviceCtx* dc = ctx->devs[di];
    CU(cudaSetDevice(dc->dev));
    ModelDev& md = m->devs[di];
    Workspace* ws = nullptr;
    rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    struct Rel { DeviceCtx* d; Workspace* w; ~Rel() { ws_release(d, w); } } rel{dc, ws};
    cudaStream_t st = ws->st;
    const size_t na = axis.size();
    const size_t chunk = (size_t)std::min<unsigned long long>(g_end - g_begin, 1ull << 21);
    // io: q (3) | f (1) | selected f (1) | selected index (1, as 8-byte integers) | axis | counter
    rc = ws_reserve(&ws->io, &ws->io_cap, 14 * std::max(chunk + na + 8, (size_t)TB));
    if (rc) return rc;
    const size_t cap = ws->io_cap / 14;
    double* dq = ws->io; double* df = dq + 3 * cap; double* dself = df + cap;
    unsigned long long* dselidx = reinterpret_cast<unsigned long long*>(dself + cap);
    double* daxis = dself + 2 * cap;
    unsigned int* dcounter = reinterpret_cast<unsigned int*>(daxis + na);
    CU(cudaMemcpyAsync(daxis, axis.data(), na * sizeof(double), cudaMemcpyHostToDevice, st));
    const size_t N = m->N, ld = m->cap;
    std::vector<unsigned long long> hidx;
    std::vector<double> hf;
    CU(cudaEventRecord(ws->ev[0], st));
    for (unsigned long long g0 = g_begin; g0 < g_end; g0 += chunk) {
        const int cnt = (int)std::min<unsigned long long>(chunk, g_end - g0);
        CU(cudaMemsetAsync(dcounter, 0, sizeof(unsigned int), st));
        CU(launch_grid_fill(daxis, (int)na, g0, cnt, dq, dq + cap, dq + 2 * cap, st));
        const bool warp_mode = (size_t)cnt <= (size_t)64 * dc->num_sms;
        int nsplit = 1;
        if (!warp_mode) {
            nsplit = predict_split(cnt, (int)N, dc->num_sms);
            if (nsplit > 1) { rc = ws_reserve(&ws->mpart, &ws->mpart_dbl, predict_part_doubles(cnt, (int)N)); if (rc) return rc; }
        }
        CU(launch_predict(md.xyz, md.xyz + ld, md.xyz + 2 * ld, md.alpha, (int)m->n, (int)N, dq, dq + cap, dq + 2 * cap, cnt,
                          df, nullptr, 0, nullptr, 0, (int)m->n_spd, m->kp, warp_mode, ws->mpart, nsplit, st));
        CU(launch_grid_select(df, g0, cnt, tol, dcounter, dselidx, dself, st));
        unsigned int found = 0;
        CU(cudaMemcpyAsync(&found, dcounter, sizeof(found), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (found > (unsigned)cnt) {
            char b[160];
            snprintf(b, sizeof b, "iso-surface sampler: selection counter %u exceeds the chunk size %d", found, cnt);
            return fail(GPR_ERR_CUDA, b);
        }
        if (found) {
            hidx.resize(found); hf.resize(found);
            CU(cudaMemcpyAsync(hidx.data(), dselidx, found * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(hf.data(), dself, found * sizeof(double), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (unsigned int i = 0; i < found; ++i) hits.emplace_back(hidx[i], hf[i]);
        }
    }
    CU(cudaEventRecord(ws->ev[1], st));
    CU(cudaStreamSynchronize(st));
    *ms = ev_ms(ws->ev[0], ws->ev[1]);
    return GPR_OK;
}

int gpr_sample_isosurface(gpr_ctx* ctx, gpr_model* m, double lo, double hi, double step, double tol, size_t capacity,
                          double* x, double* y, double* z, double* f, double* var, size_t* count) {
    if (!ctx) return fail(GPR_ERR_INVALID, "null context");
    if (!m) return fail(GPR_ERR_INVALID, "Empty Model pointer");
    if (!count || !(step > 0.0) || !(hi >= lo) || !(tol >= 0.0)) return fail(GPR_ERR_INVALID, "bad lattice or null count");
    // the lattice axis, accumulated exactly like the node's loops: for (x = -scale; x <= scale; x += pass)
    std::vector<double> axis;
    for (double a = lo; a <= hi; a += step) {
        axis.push_back(a);
        if (axis.size() > 4096) return fail(GPR_ERR_INVALID, "lattice has more than 4096 points per axis");
    }
    const size_t na = axis.size();
    const unsigned long long total = (unsigned long long)na * na * na;
    // contiguous lattice ranges over the context's devices (each query is independent: the result does not depend
    // on the split); small lattices stay on the primary device
    const size_t nd = ctx->devs.size();
    const size_t use = (nd > 1 && total >= ((unsigned long long)nd << 20)) ? nd : 1;
    std::vector<std::vector<std::pair<unsigned long long, double>>> part(use);
    std::vector<int> rcs(use, 0);
    std::vector<std::string> errs(use);
    std::vector<double> tms(use, 0.0);
    auto work = [&](size_t di) {
        rcs[di] = sample_range(m, di, axis, total * di / use, total * (di + 1) / use, tol, part[di], &tms[di]);
        if (rcs[di]) errs[di] = g_err;
    };
    if (use == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (size_t di = 0; di < use; ++di) pool.emplace_back(work, di);
        for (auto& t : pool) t.join();
    }
    for (size_t di = 0; di < use; ++di) if (rcs[di]) return fail(rcs[di], errs[di]);
    std::vector<std::pair<unsigned long long, double>> hits;
    for (size_t di = 0; di < use; ++di) hits.insert(hits.end(), part[di].begin(), part[di].end());
    {
        std::lock_guard<std::mutex> lk(ctx->tmu);
        ctx->timings.predict_mean_ms = *std::max_element(tms.begin(), tms.end());
    }
    std::sort(hits.begin(), hits.end());                 // lattice order: deterministic output
    *count = hits.size();
    const size_t keep = std::min(hits.size(), capacity);
    if (keep == 0 || (!x && !y && !z && !f && !var)) return GPR_OK;
    std::vector<double> sx(keep), sy(keep), sz(keep), sf(keep);
    for (size_t i = 0; i < keep; ++i) {
        const unsigned long long g = hits[i].first;
        sx[i] = axis[(size_t)(g / ((unsigned long long)na * na))];
        sy[i] = axis[(size_t)((g / na) % na)];
        sz[i] = axis[(size_t)(g % na)];
        sf[i] = hits[i].second;
    }
    if (var) {
        // the expensive part (n^2 flop per point) only for the survivors, through the regular predict path
        std::vector<double> f2(keep);
        const double mean_ms = ctx->timings.predict_mean_ms;
        int rc = gpr_predict(ctx, m, sx.data(), sy.data(), sz.data(), keep, f2.data(), var, nullptr, nullptr, nullptr);
        if (rc) return rc;
        std::lock_guard<std::mutex> lk(ctx->tmu);
        ctx->timings.predict_mean_ms += mean_ms;
        ctx->timings.predict_total_ms += mean_ms;
    }
    if (x) memcpy(x, sx.data(), keep * sizeof(double));
    if (y) memcpy(y, sy.data(), keep * sizeof(double));
    if (z) memcpy(z, sz.data(), keep * sizeof(double));
    if (f) memcpy(f, sf.data(), keep * sizeof(double));
    return GPR_OK;
}

}  // extern "C"

// Points [a, b) of a batched projection on device slot di.
static int project_range(gpr_model* m, size_t di, const double* const src[6], size_t a, size_t b, double f_tol, double improve_tol,
                         unsigned max_iter, double step_mul, double* ox, double* oy, double* oz, int* hstat, double* ms) {
    gpr_ctx* ctx = m->ctx;
    int rc = ensure_on_device(m, di, false);
    if (rc) return rc;
    DeviceCtx* dc = ctx->devs[di];
    CU(cudaSetDevice(dc->dev));
    ModelDev& md = m->devs[di];
    Workspace* ws = nullptr;
    rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    struct Rel { DeviceCtx* d; Workspace* w; ~Rel() { ws_release(d, w); } } rel{dc, ws};
    cudaStream_t st = ws->st;
    const size_t count = b - a;
    rc = ws_reserve(&ws->io, &ws->io_cap, 14 * std::max(count, (size_t)TB));
    if (rc) return rc;
    const size_t cap = ws->io_cap / 14, ld = m->cap;
    double* din = ws->io;                  // x|y|z|nx|ny|nz
    double* dout = din + 6 * cap;          // x|y|z
    int* dstat = reinterpret_cast<int*>(dout + 3 * cap);
    for (int c = 0; c < 6; ++c) CU(cudaMemcpyAsync(din + c * cap, src[c] + a, count * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(ws->ev[0], st));
    CU(launch_project(md.xyz, md.xyz + ld, md.xyz + 2 * ld, md.alpha, (int)m->n, din, cap, (int)count, f_tol, improve_tol,
                      (int)max_iter, step_mul, dout, dstat, m->kp, st));
    CU(cudaEventRecord(ws->ev[1], st));
    CU(cudaMemcpyAsync(ox + a, dout, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(oy + a, dout + cap, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(oz + a, dout + 2 * cap, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hstat + a, dstat, count * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *ms = ev_ms(ws->ev[0], ws->ev[1]);
    return GPR_OK;
}

extern "C" {

int gpr_project(gpr_ctx* ctx, gpr_model* m, const double* x, const double* y, const double* z, const double* nx,
                const double* ny, const double* nz, size_t count, double f_tol, double improve_tol, unsigned max_iter,
                double step_mul, double* ox, double* oy, double* oz, int* status) {
    if (!ctx) return fail(GPR_ERR_INVALID, "null context");
    if (!m) return fail(GPR_ERR_INVALID, "Empty Model pointer");
    if (!x || !y || !z || !nx || !ny || !nz || !ox || !oy || !oz || count == 0) return fail(GPR_ERR_INVALID, "All input data is empty!");
    if (count > (size_t)1 << 24 || max_iter > (unsigned)1 << 24) return fail(GPR_ERR_INVALID, "too many points or iterations");
    const double* src[6] = {x, y, z, nx, ny, nz};
    std::vector<int> hstat(count);
    // every point is an independent iteration (one CTA each): contiguous ranges over the context's devices once there
    // are enough points to fill them; the result does not depend on the split
    const size_t nd = ctx->devs.size();
    const size_t use = (nd > 1 && count >= 64 * nd) ? nd : 1;
    std::vector<int> rcs(use, 0);
    std::vector<std::string> errs(use);
    std::vector<double> tms(use, 0.0);
    auto work = [&](size_t di) {
        rcs[di] = project_range(m, di, src, count * di / use, count * (di + 1) / use, f_tol, improve_tol, max_iter, step_mul,
                                ox, oy, oz, hstat.data(), &tms[di]);
        if (rcs[di]) errs[di] = g_err;
    };
    if (use == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (size_t di = 0; di < use; ++di) pool.emplace_back(work, di);
        for (auto& t : pool) t.join();
    }
    for (size_t di = 0; di < use; ++di) if (rcs[di]) return fail(rcs[di], errs[di]);
    {
        std::lock_guard<std::mutex> lk(ctx->tmu);
        ctx->timings.predict_mean_ms = *std::max_element(tms.begin(), tms.end());
        ctx->timings.predict_total_ms = ctx->timings.predict_mean_ms;
    }
    bool bad = false;
    for (size_t i = 0; i < count; ++i) { if (status) status[i] = hstat[i]; bad = bad || hstat[i] == INT_MIN; }
    if (bad) return fail(GPR_ERR_INVALID, "f is nan or inf");       // include/atlas/atlas.hpp:230
    return GPR_OK;
}

// Batched counterpart of the node's marchingSampling / marchingCubes (src/gp_node.cpp:1103-1190, :1195-1292): a flood fill
// over cubes of edge `leaf` that follows the iso-surface.  The reference runs one std::thread per cube, recursively, and
// one evaluate(q = 1) per cube sample; here every WAVE of the flood fill (all frontier cubes) is one batched mean
// evaluation, followed by one batched variance evaluation of the wave's kept samples.  Coordinates follow the
// reference's arithmetic exactly: the start point and the cube centres are pcl::PointXYZ (float), a sample coordinate
// is the float expression start.x - leaf/2 + i*pass widened to double (:1209-1211).
int gpr_sample_marching(gpr_ctx* ctx, gpr_model* m, double grid_lo, double grid_hi, double grid_step, float leaf, float pass,
                        double tol, size_t capacity, double* x, double* y, double* z, double* f, double* var, size_t* count,
                        size_t* cubes) {
    if (!ctx) return fail(GPR_ERR_INVALID, "null context");
    if (!m) return fail(GPR_ERR_INVALID, "Empty Model pointer");
    if (!count || !(grid_step > 0.0) || !(grid_hi >= grid_lo) || !(leaf > 0.0f) || !(pass > 0.0f) || !(tol >= 0.0))
        return fail(GPR_ERR_INVALID, "bad lattice, cube size or null count");
    *count = 0;
    if (cubes) *cubes = 0;
    // 1. the starting point: first lattice point, in the reference's scan order (x outermost), with |f| <= tol (:1124-1150)
    std::vector<double> axis;
    for (double a = grid_lo; a <= grid_hi; a += grid_step) {
        axis.push_back(a);
        if (axis.size() > 1024) return fail(GPR_ERR_INVALID, "start lattice has more than 1024 points per axis");
    }
    const size_t na = axis.size(), nl = na * na * na;
    std::vector<double> lx(nl), ly(nl), lz(nl), lf(nl);
    for (size_t i = 0, g = 0; i < na; ++i)
        for (size_t j = 0; j < na; ++j)
            for (size_t k = 0; k < na; ++k, ++g) { lx[g] = axis[i]; ly[g] = axis[j]; lz[g] = axis[k]; }
    int rc = predict_host(ctx, m, lx.data(), ly.data(), lz.data(), nl, lf.data(), nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    size_t s0 = nl;
    for (size_t g = 0; g < nl; ++g) if (std::fabs(lf[g]) <= tol) { s0 = g; break; }
    if (s0 == nl) return fail(GPR_ERR_INVALID, "No starting point found. Relax grid pass.");          // :1153
    const float sx = (float)lx[s0], sy = (float)ly[s0], sz = (float)lz[s0];
    const long steps = std::lround(leaf / pass);                                                          // :1201
    if (steps < 1 || steps > 64) return fail(GPR_ERR_INVALID, "leaf / pass must round to 1 .. 64");
    const size_t per_cube = (size_t)(steps + 1) * (steps + 1) * (steps + 1);
    // 2. flood fill by waves.  A cube is identified by its integer offset from the start cube; its centre is obtained
    // like the reference's recursion does, by repeated float additions of +-leaf along the path — which is the same
    // value for every path only up to float rounding, so the centre is computed canonically: x first, then y, then z.
    typedef std::array<long, 3> Idx;
    auto centre = [&](const Idx& c, float out[3]) {
        const float s[3] = {sx, sy, sz};
        for (int a = 0; a < 3; ++a) {
            float v = s[a];
            for (long t = 0; t < std::labs(c[a]); ++t) v = c[a] > 0 ? v + leaf : v - leaf;
            out[a] = v;
        }
    };
    std::set<Idx> visited;
    std::set<Idx> emitted;                       // lattice index of a kept sample in units of `pass`: shared faces are emitted once
    std::vector<Idx> frontier(1, Idx{{0, 0, 0}});
    visited.insert(frontier[0]);
    std::vector<double> ox, oy, oz, of;
    size_t ncubes = 0;
    const size_t MAX_CUBES = (size_t)1 << 20;
    while (!frontier.empty()) {
        std::sort(frontier.begin(), frontier.end());
        const size_t nc = frontier.size();
        ncubes += nc;
        if (ncubes > MAX_CUBES) return fail(GPR_ERR_INVALID, "marching sampler: more than 2^20 cubes");
        std::vector<double> qx(nc * per_cube), qy(nc * per_cube), qz(nc * per_cube), qf(nc * per_cube);
        for (size_t c = 0; c < nc; ++c) {
            float ctr[3];
            centre(frontier[c], ctr);
            size_t p = c * per_cube;
            for (long i = 0; i <= steps; ++i)
                for (long j = 0; j <= steps; ++j)
                    for (long k = 0; k <= steps; ++k, ++p) {
                        qx[p] = (double)(ctr[0] - leaf / 2 + (float)i * pass);                             // :1209-1211
                        qy[p] = (double)(ctr[1] - leaf / 2 + (float)j * pass);
                        qz[p] = (double)(ctr[2] - leaf / 2 + (float)k * pass);
                    }
        }
        rc = predict_host(ctx, m, qx.data(), qy.data(), qz.data(), qx.size(), qf.data(), nullptr, nullptr, nullptr, nullptr);
        if (rc) return rc;
        std::vector<Idx> next;
        for (size_t c = 0; c < nc; ++c) {
            bool where[6] = {false, false, false, false, false, false};
            size_t p = c * per_cube;
            const Idx& ci = frontier[c];
            for (long i = 0; i <= steps; ++i)
                for (long j = 0; j <= steps; ++j)
                    for (long k = 0; k <= steps; ++k, ++p) {
                        if (!(std::fabs(qf[p]) <= tol)) continue;                                        // :1218
                        if (i == 0) where[0] = true;
                        if (i == steps) where[1] = true;
                        if (j == 0) where[2] = true;
                        if (j == steps) where[3] = true;
                        if (k == 0) where[4] = true;
                        if (k == steps) where[5] = true;
                        const Idx li{{ci[0] * steps + i, ci[1] * steps + j, ci[2] * steps + k}};
                        if (emitted.insert(li).second) { ox.push_back(qx[p]); oy.push_back(qy[p]); oz.push_back(qz[p]); of.push_back(qf[p]); }
                    }
            for (int d = 0; d < 6; ++d) {                                                                 // :1262-1288
                if (!where[d]) continue;
                Idx nb = ci;
                nb[d / 2] += (d % 2) ? 1 : -1;
                if (visited.insert(nb).second) next.push_back(nb);
            }
        }
        frontier.swap(next);
    }
    *count = ox.size();
    if (cubes) *cubes = ncubes;
    const size_t keep = std::min(ox.size(), capacity);
    if (keep == 0) return GPR_OK;
    if (var) {
        std::vector<double> f2(keep);
        rc = predict_host(ctx, m, ox.data(), oy.data(), oz.data(), keep, f2.data(), var, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    if (x) memcpy(x, ox.data(), keep * sizeof(double));
    if (y) memcpy(y, oy.data(), keep * sizeof(double));
    if (z) memcpy(z, oz.data(), keep * sizeof(double));
    if (f) memcpy(f, of.data(), keep * sizeof(double));
    return GPR_OK;
}

int gpr_sample_chart(gpr_ctx* ctx, gpr_model* m, const double* frames, const size_t* counts, size_t n_charts,
                     const double* r_in, const double* th_in, unsigned long long seed, double* sx, double* sy, double* sz,
                     double* f, double* v, size_t* order) {
    if (!ctx) return fail(GPR_ERR_INVALID, "null context");
    if (!m) return fail(GPR_ERR_INVALID, "Empty Model pointer");
    if (!frames || !counts || n_charts == 0 || (r_in == nullptr) != (th_in == nullptr)) return fail(GPR_ERR_INVALID, "All input data is empty!");
    if (n_charts > (size_t)1 << 20) return fail(GPR_ERR_INVALID, "too many charts");
    std::vector<unsigned long long> offs(n_charts + 1, 0);
    for (size_t c = 0; c < n_charts; ++c) {
        if (counts[c] > (size_t)1 << 20) return fail(GPR_ERR_INVALID, "too many samples on one chart");
        offs[c + 1] = offs[c] + counts[c];
    }
    const size_t total = (size_t)offs[n_charts];
    if (total == 0) return GPR_OK;
    if (total > (size_t)1 << 26) return fail(GPR_ERR_INVALID, "too many samples");
    std::vector<double> rr, tt;
    if (!r_in) {
        // the reference's generator and call order (include/random_generation.hpp:9-27; atlas_variance.hpp:176-177: r then th
        // per sample), seeded by the caller instead of std::random_device
        std::mt19937_64 eng(seed);
        rr.resize(total); tt.resize(total);
        std::uniform_real_distribution<double> dr(0.8, std::nextafter(1.0, std::numeric_limits<double>::max())), dt(0.0, 2 * M_PI);
        for (size_t i = 0; i < total; ++i) { rr[i] = dr(eng); tt[i] = dt(eng); }
        r_in = rr.data(); th_in = tt.data();
    }
    int rc = ensure_on_device(m, 0, false);
    if (rc) return rc;
    DeviceCtx* dc = ctx->devs[0];
    CU(cudaSetDevice(dc->dev));
    Workspace* ws = nullptr;
    rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    struct Rel { DeviceCtx* d; Workspace* w; ~Rel() { ws_release(d, w); } } rel{dc, ws};
    cudaStream_t st = ws->st;
    // io: q (3 x total) | f | v | r | th | order (8-byte ints) | frames (13 n) | offsets (n + 1) | bad flag
    const size_t need = 8 * total + 13 * n_charts + (n_charts + 1) + 2;
    rc = ws_reserve(&ws->io, &ws->io_cap, std::max(need, (size_t)14 * TB));
    if (rc) return rc;
    double* dq = ws->io; double* df = dq + 3 * total; double* dv = df + total; double* dr_ = dv + total; double* dth = dr_ + total;
    unsigned long long* dord = reinterpret_cast<unsigned long long*>(dth + total);
    double* dfr = dth + 2 * total;
    unsigned long long* doff = reinterpret_cast<unsigned long long*>(dfr + 13 * n_charts);
    int* dbad = reinterpret_cast<int*>(doff + n_charts + 1);
    CU(cudaMemcpyAsync(dfr, frames, 13 * n_charts * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(doff, offs.data(), (n_charts + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dr_, r_in, total * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dth, th_in, total * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dbad, 0, sizeof(int), st));
    CU(launch_chart_fill(dfr, doff, (int)n_charts, dr_, dth, (int)total, dq, dq + total, dq + 2 * total, st));
    CU(cudaStreamSynchronize(st));
    PredictIO io;
    io.qx = dq; io.qy = dq + total; io.qz = dq + 2 * total; io.q = total; io.offset = 0;
    io.f = df; io.var = dv; io.grad = nullptr; io.tx = nullptr; io.ty = nullptr; io.out_ld = total; io.device_ptrs = true;
    double tm = 0, tv = 0, th2 = 0, td = 0;
    rc = predict_on_device(m, 0, io, &tm, &tv, &th2, &td);
    if (rc) return rc;
    CU(cudaSetDevice(dc->dev));
    CU(launch_chart_rank(df, dv, doff, (int)n_charts, dord, dbad, st));
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (sx) CU(cudaMemcpyAsync(sx, dq, total * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sy) CU(cudaMemcpyAsync(sy, dq + total, total * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sz) CU(cudaMemcpyAsync(sz, dq + 2 * total, total * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (f) CU(cudaMemcpyAsync(f, df, total * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (v) CU(cudaMemcpyAsync(v, dv, total * sizeof(double), cudaMemcpyDeviceToHost, st));
    std::vector<unsigned long long> hord(order ? total : 0);
    if (order) CU(cudaMemcpyAsync(hord.data(), dord, total * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (order) for (size_t i = 0; i < total; ++i) order[i] = (size_t)hord[i];
    {
        std::lock_guard<std::mutex> lk(ctx->tmu);
        ctx->timings.predict_mean_ms = tm; ctx->timings.predict_var_ms = tv; ctx->timings.predict_total_ms = tm + tv;
    }
    if (bad) return fail(GPR_ERR_INVALID, "v is nan or inf");          // include/atlas/atlas_variance.hpp:210
    return GPR_OK;
}

// ---- model export / import (SURVEY §5 "checkpoint / resume", §8(f).4) ---------------------------------------
// File: header | x y z label [sigma2] alpha [normals] | [lower triangle of L by columns, n(n+1)/2 doubles].
struct SaveHeader {
    char magic[8];               // "GPRB200\0"
    unsigned int version, kind;
    double p0, p1, R;
    unsigned long long n, n_normals, n_tail;
    unsigned int has_s2, with_normals, has_factor, reserved;
};

int gpr_model_save(gpr_ctx* ctx, gpr_model* m, const char* path, int with_factor) {
    if (!ctx || !m || !path) return fail(GPR_ERR_INVALID, "null pointer");
    if (m->replica) return fail(GPR_ERR_INVALID, "cannot save a replica model");
    std::lock_guard<std::mutex> lk(m->mu);
    FILE* fh = fopen(path, "wb");
    if (!fh) return fail(GPR_ERR_INVALID, std::string("cannot open ") + path + " for writing");
    struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{fh};
    SaveHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "GPRB200", 8);
    h.version = 1; h.kind = (unsigned)m->kernel.kind; h.p0 = m->kernel.p0; h.p1 = m->kernel.p1; h.R = m->R;
    h.n = m->n; h.n_normals = m->n_normals; h.n_tail = m->n_tail;
    h.has_s2 = m->has_s2; h.with_normals = m->with_normals;
    // a tail model is refitted on load; so is one whose internal point order differs from the caller's (L is in
    // internal order, the stored training set in the caller's)
    h.has_factor = (with_factor && m->n_tail == 0 && m->perm.empty()) ? 1 : 0;
    bool ok = fwrite(&h, sizeof h, 1, fh) == 1;
    auto put = [&](const std::vector<double>& v) { if (!v.empty()) ok = ok && fwrite(v.data(), sizeof(double), v.size(), fh) == v.size(); };
    put(m->hx); put(m->hy); put(m->hz); put(m->hlabel);
    if (m->has_s2) put(m->hs2);
    put(m->h_alpha);
    put(m->h_normals);
    if (h.has_factor) {
        CU(cudaSetDevice(m->devs[0].dev));
        const size_t n = m->n;
        std::vector<double> blk(n * TB), packed;
        for (size_t c0 = 0; c0 < n; c0 += TB) {
            const size_t cols = std::min<size_t>(TB, n - c0);
            CU(cudaMemcpy2D(blk.data(), n * sizeof(double), m->L + c0 * m->cap, m->cap * sizeof(double), n * sizeof(double), cols,
                            cudaMemcpyDeviceToHost));
            packed.clear();
            for (size_t c = 0; c < cols; ++c) packed.insert(packed.end(), blk.begin() + c * n + (c0 + c), blk.begin() + (c + 1) * n);
            ok = ok && fwrite(packed.data(), sizeof(double), packed.size(), fh) == packed.size();
        }
    }
    if (!ok) return fail(GPR_ERR_INVALID, std::string("short write to ") + path);
    return GPR_OK;
}

int gpr_model_load(gpr_ctx* ctx, const char* path, gpr_model** out) {
    if (!ctx || !path || !out) return fail(GPR_ERR_INVALID, "null pointer");
    FILE* fh = fopen(path, "rb");
    if (!fh) return fail(GPR_ERR_INVALID, std::string("cannot open ") + path);
    struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{fh};
    SaveHeader h;
    if (fread(&h, sizeof h, 1, fh) != 1 || memcmp(h.magic, "GPRB200", 8) != 0 || h.version != 1 || h.kind > 2 || h.n == 0)
        return fail(GPR_ERR_INVALID, std::string(path) + " is not a GPRB200 model file");
    // The header is untrusted input: every count must be consistent with the others and with the file size before
    // anything is allocated from it.
    {
        const unsigned long long n64 = h.n;
        if (n64 > (1ull << 24) || (h.n_normals != 0 && h.n_normals != n64) || h.n_tail >= n64 || h.has_s2 > 1 ||
            h.with_normals > 1 || h.has_factor > 1 || (h.has_factor && h.n_tail != 0))
            return fail(GPR_ERR_INVALID, std::string(path) + " has an inconsistent header");
        unsigned long long doubles = (5ull + h.has_s2) * n64 + 3ull * h.n_normals;
        if (h.has_factor) doubles += n64 * (n64 + 1) / 2;
        long long fsize = -1;
        if (fseek(fh, 0, SEEK_END) == 0) fsize = ftell(fh);
        if (fsize < 0 || fseek(fh, (long)sizeof h, SEEK_SET) != 0 || (unsigned long long)fsize != sizeof h + 8ull * doubles)
            return fail(GPR_ERR_INVALID, std::string(path) + " is truncated or has trailing data (size does not match its header)");
    }
    gpr_model* m = nullptr;
    try {
    m = new gpr_model();
    auto bail = [&](int rc) { free_factor(m); delete m; return rc; };
    m->ctx = ctx; m->kernel = gpr_kernel_t{(int)h.kind, h.p0, h.p1}; m->kp = make_kp(m->kernel); m->k0 = kernel_at_zero(m->kp);
    m->has_s2 = h.has_s2 != 0; m->with_normals = h.with_normals != 0;
    const size_t n = (size_t)h.n;
    bool ok = true;
    auto get = [&](std::vector<double>& v, size_t cnt) { v.resize(cnt); if (cnt) ok = ok && fread(v.data(), sizeof(double), cnt, fh) == cnt; };
    std::vector<double> alpha, normals;
    get(m->hx, n); get(m->hy, n); get(m->hz, n); get(m->hlabel, n);
    if (m->has_s2) get(m->hs2, n);
    get(alpha, n);
    get(normals, 3 * (size_t)h.n_normals);
    if (!ok) return bail(fail(GPR_ERR_INVALID, std::string(path) + " is truncated"));
    if (!h.has_factor) {
        // no factor in the file (or an indefinite-tail model): refit from the stored training set
        const bool wn = m->with_normals;
        m->with_normals = false;
        int rc = fit_from_host(m, false);
        m->with_normals = wn;
        if (rc) return bail(rc);
        m->R = h.R; m->h_normals = normals; m->n_normals = (size_t)h.n_normals;
        *out = m;
        return GPR_OK;
    }
    DeviceCtx* dc = ctx->devs[0];
    if (cudaSetDevice(dc->dev) != cudaSuccess) return bail(fail(GPR_ERR_CUDA, "cudaSetDevice failed"));
    const size_t N = (n + TB - 1) / TB * TB;
    const int nb = (int)(N / TB);
    m->n = n; m->N = N; m->nb = nb; m->cap = N; m->n_spd = n; m->n_tail = 0; m->mp = 0; m->R = h.R;
    m->devs.assign(ctx->devs.size(), ModelDev());
    for (size_t i = 0; i < ctx->devs.size(); ++i) m->devs[i].dev = ctx->devs[i]->dev;
    ModelDev& md = m->devs[0];
#define LOAD_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(fail(e__ == cudaErrorMemoryAllocation ? GPR_ERR_OOM : GPR_ERR_CUDA, std::string("gpr_model_load: ") + cudaGetErrorString(e__))); } while (0)
    LOAD_TRY(cudaMalloc((void**)&md.xyz, 3 * N * sizeof(double)));
    LOAD_TRY(cudaMalloc((void**)&md.alpha, N * sizeof(double)));
    LOAD_TRY(cudaMalloc((void**)&m->label, N * sizeof(double)));
    LOAD_TRY(cudaMalloc((void**)&m->s2, N * sizeof(double)));
    LOAD_TRY(cudaMalloc((void**)&m->zfwd, N * sizeof(double)));
    LOAD_TRY(cudaMalloc((void**)&m->L, N * N * sizeof(double)));
    LOAD_TRY(cudaMalloc((void**)&m->Dinv, (size_t)nb * TB * TB * sizeof(double)));
    LOAD_TRY(cudaMalloc((void**)&m->scratch, (8 + (size_t)nb * nb) * sizeof(int)));
    LOAD_TRY(cudaMemset(md.xyz, 0, 3 * N * sizeof(double)));
    LOAD_TRY(cudaMemset(md.alpha, 0, N * sizeof(double)));
    LOAD_TRY(cudaMemset(m->label, 0, N * sizeof(double)));
    LOAD_TRY(cudaMemset(m->s2, 0, N * sizeof(double)));
    LOAD_TRY(cudaMemcpy(md.xyz, m->hx.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    LOAD_TRY(cudaMemcpy(md.xyz + N, m->hy.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    LOAD_TRY(cudaMemcpy(md.xyz + 2 * N, m->hz.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    LOAD_TRY(cudaMemcpy(m->label, m->hlabel.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    if (m->has_s2) LOAD_TRY(cudaMemcpy(m->s2, m->hs2.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    LOAD_TRY(cudaMemcpy(md.alpha, alpha.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    // the factor: 128 columns at a time, padded rows / columns are identity
    std::vector<double> blk(N * TB), packed;
    for (size_t c0 = 0; c0 < N; c0 += TB) {
        std::fill(blk.begin(), blk.end(), 0.0);
        for (size_t c = 0; c < TB; ++c) {
            const size_t gc = c0 + c;
            if (gc < n) {
                packed.resize(n - gc);
                if (fread(packed.data(), sizeof(double), n - gc, fh) != n - gc) return bail(fail(GPR_ERR_INVALID, std::string(path) + " is truncated"));
                std::copy(packed.begin(), packed.end(), blk.begin() + c * N + gc);
            } else {
                blk[c * N + gc] = 1.0;
            }
        }
        LOAD_TRY(cudaMemcpy(m->L + c0 * N, blk.data(), N * TB * sizeof(double), cudaMemcpyHostToDevice));
    }
    LOAD_TRY(launch_dinv_from_l(m->L, N, nb, m->Dinv, 0));
    LOAD_TRY(cudaDeviceSynchronize());
#undef LOAD_TRY
    m->h_alpha = alpha; m->h_normals = normals; m->n_normals = (size_t)h.n_normals;
    md.have = true;
    md.lfac = m->L; md.dinv = m->Dinv; md.own_fac = false; md.have_fac = true;
    *out = m;
    return GPR_OK;
    } catch (const std::bad_alloc&) {
        if (m) { free_factor(m); delete m; }
        return fail(GPR_ERR_OOM, "gpr_model_load: out of host memory");
    } catch (const std::exception& e) {
        if (m) { free_factor(m); delete m; }
        return fail(GPR_ERR_INVALID, std::string("gpr_model_load: ") + e.what());
    }
}

int gpr_model_prepare_variance(gpr_ctx* ctx, gpr_model* m) {
    if (!ctx || !m) return fail(GPR_ERR_INVALID, "Empty Model pointer");
    for (size_t di = 0; di < ctx->devs.size(); ++di) {
        int rc = ensure_on_device(m, di, true);
        if (rc) return rc;
    }
    return GPR_OK;
}

int gpr_model_solve(gpr_ctx* ctx, gpr_model* m, const double* B, size_t nrhs, double* X) {
    if (!ctx || !m) return fail(GPR_ERR_INVALID, "Empty Model pointer");
    if (!B || !X || nrhs == 0) return fail(GPR_ERR_INVALID, "All input data is empty!");
    if (m->replica || !m->L) return fail(GPR_ERR_INVALID, "model holds no factor (replica)");
    if (m->n_tail > 0 || !m->perm.empty()) return fail(GPR_ERR_INVALID, "gpr_model_solve needs a positive definite model in the caller's point order");
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceCtx* dc = ctx->devs[0];
    CU(cudaSetDevice(dc->dev));
    Workspace* ws = nullptr;
    int rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    struct Rel { DeviceCtx* d; Workspace* w; ~Rel() { ws_release(d, w); } } rel{dc, ws};
    cudaStream_t st = ws->st;
    const size_t N = m->N, ld = m->cap, n = m->n;
    rc = ws_reserve(&ws->mpart, &ws->mpart_dbl, 2 * N);
    if (rc) return rc;
    double* rhs = ws->mpart; double* sol = rhs + N;
    for (size_t c = 0; c < nrhs; ++c) {
        CU(cudaMemsetAsync(rhs, 0, N * sizeof(double), st));
        CU(cudaMemcpyAsync(rhs, B + c * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
        CU(launch_trsv(0, m->L, ld, m->nb, m->Dinv, rhs, m->zfwd, m->scratch, dc->num_sms, st));
        CU(launch_trsv(1, m->L, ld, m->nb, m->Dinv, m->zfwd, sol, m->scratch, dc->num_sms, st));
        CU(cudaMemcpyAsync(X + c * n, sol, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    int flags[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(flags, m->scratch, sizeof(flags), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (flags[2] != 0) return fail(GPR_ERR_CUDA, "triangular solve kernel aborted (dependency wait timed out)");
    return GPR_OK;
}

// Reallocate every per-model buffer of the primary device with leading dimension newcap (multiple of 128,
// >= m->N), keeping the fitted state; rows [N, newcap) of L and L^-1 become identity padding.
static int grow_capacity(gpr_model* m, size_t newcap, cudaStream_t st) {
    if (newcap <= m->cap) return GPR_OK;
    ModelDev& md = m->devs[0];
    const size_t N = m->N, oc = m->cap;
    const int nbc = (int)(newcap / TB);
    double *xyz = nullptr, *alpha = nullptr, *label = nullptr, *s2 = nullptr, *zfwd = nullptr, *L = nullptr, *Dinv = nullptr, *X = nullptr;
    int* scratch = nullptr;
    auto drop = [&]() { cudaFree(xyz); cudaFree(alpha); cudaFree(label); cudaFree(s2); cudaFree(zfwd); cudaFree(L); cudaFree(Dinv); cudaFree(X); cudaFree(scratch); };
#define GROW_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { drop(); return fail(e__ == cudaErrorMemoryAllocation ? GPR_ERR_OOM : GPR_ERR_CUDA, std::string("grow_capacity: ") + cudaGetErrorString(e__)); } } while (0)
    GROW_TRY(cudaMalloc((void**)&xyz, 3 * newcap * sizeof(double)));
    GROW_TRY(cudaMalloc((void**)&alpha, newcap * sizeof(double)));
    GROW_TRY(cudaMalloc((void**)&label, newcap * sizeof(double)));
    GROW_TRY(cudaMalloc((void**)&s2, newcap * sizeof(double)));
    GROW_TRY(cudaMalloc((void**)&zfwd, newcap * sizeof(double)));
    GROW_TRY(cudaMalloc((void**)&L, newcap * newcap * sizeof(double)));
    GROW_TRY(cudaMalloc((void**)&Dinv, (size_t)nbc * TB * TB * sizeof(double)));
    GROW_TRY(cudaMalloc((void**)&scratch, (8 + (size_t)nbc * nbc) * sizeof(int)));
    if (md.have_linv) GROW_TRY(cudaMalloc((void**)&X, newcap * newcap * sizeof(double)));
    GROW_TRY(cudaMemsetAsync(xyz, 0, 3 * newcap * sizeof(double), st));
    GROW_TRY(cudaMemsetAsync(alpha, 0, newcap * sizeof(double), st));
    GROW_TRY(cudaMemsetAsync(label, 0, newcap * sizeof(double), st));
    GROW_TRY(cudaMemsetAsync(s2, 0, newcap * sizeof(double), st));
    GROW_TRY(cudaMemsetAsync(zfwd, 0, newcap * sizeof(double), st));
    for (int c = 0; c < 3; ++c)
        GROW_TRY(cudaMemcpyAsync(xyz + c * newcap, md.xyz + c * oc, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    GROW_TRY(cudaMemcpyAsync(alpha, md.alpha, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    GROW_TRY(cudaMemcpyAsync(label, m->label, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    GROW_TRY(cudaMemcpyAsync(s2, m->s2, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    GROW_TRY(cudaMemcpy2DAsync(L, newcap * sizeof(double), m->L, oc * sizeof(double), N * sizeof(double), N, cudaMemcpyDeviceToDevice, st));
    if (X) GROW_TRY(cudaMemcpy2DAsync(X, newcap * sizeof(double), md.linv, oc * sizeof(double), N * sizeof(double), N, cudaMemcpyDeviceToDevice, st));
    GROW_TRY(cudaMemcpyAsync(Dinv, m->Dinv, (size_t)m->nb * TB * TB * sizeof(double), cudaMemcpyDeviceToDevice, st));
    GROW_TRY(launch_identity_rows(L, X, newcap, (int)N, (int)newcap, (int)newcap, st));
    // identity inverse diagonal tiles for the new tile rows (L^-1 of an identity tile)
    GROW_TRY(launch_dinv_from_x(L, newcap, m->nb, nbc - m->nb, Dinv, st));
    GROW_TRY(cudaStreamSynchronize(st));
#undef GROW_TRY
    cudaFree(md.xyz); cudaFree(md.alpha); cudaFree(m->label); cudaFree(m->s2); cudaFree(m->zfwd); cudaFree(m->L);
    cudaFree(m->Dinv); cudaFree(m->scratch); cudaFree(md.linv); cudaFree(m->aws);
    md.xyz = xyz; md.alpha = alpha; m->label = label; m->s2 = s2; m->zfwd = zfwd; m->L = L; m->Dinv = Dinv;
    m->scratch = scratch; md.linv = X; m->aws = nullptr; m->aws_dbl = 0;
    md.lfac = m->L; md.dinv = m->Dinv;
    for (size_t di = 1; di < m->devs.size(); ++di) {       // copies in the old layout
        ModelDev& d = m->devs[di];
        if (d.own_fac) { cudaSetDevice(d.dev); cudaFree(d.lfac); cudaFree(d.dinv); }
        d.lfac = d.dinv = nullptr; d.have_fac = d.own_fac = false;
    }
    cudaSetDevice(md.dev);
    m->cap = newcap;
    // replicas on the other devices have the old layout: drop them, they are re-copied on demand
    for (size_t di = 1; di < m->devs.size(); ++di) {
        ModelDev& d = m->devs[di];
        cudaSetDevice(d.dev);
        cudaFree(d.xyz); cudaFree(d.alpha); cudaFree(d.linv);
        d.xyz = d.alpha = d.linv = nullptr; d.have = d.have_linv = false;
    }
    cudaSetDevice(md.dev);
    return GPR_OK;
}

static void host_append(gpr_model* m, const double* x, const double* y, const double* z, const double* label,
                        const double* sigma2, size_t k) {
    const size_t p = m->hx.size();
    m->hx.insert(m->hx.end(), x, x + k); m->hy.insert(m->hy.end(), y, y + k); m->hz.insert(m->hz.end(), z, z + k);
    m->hlabel.insert(m->hlabel.end(), label, label + k);
    if (m->has_s2 || sigma2) {
        // gp_regressor.hpp:449-450: S2 grows with the new block; absent entries are zero noise.
        m->hs2.resize(p, 0.0);
        if (sigma2) m->hs2.insert(m->hs2.end(), sigma2, sigma2 + k); else m->hs2.resize(p + k, 0.0);
        m->has_s2 = true;
    }
}

// Incremental path of gpr_append (gpr_append.cu).  Caller holds m->mu.
static int append_incremental(gpr_model* m, const double* x, const double* y, const double* z, const double* label,
                              const double* sigma2, size_t k) {
    gpr_ctx* ctx = m->ctx;
    DeviceCtx* dc = ctx->devs[0];
    CU(cudaSetDevice(dc->dev));
    int rc = ensure_linv_primary(m);
    if (rc) return rc;
    Workspace* ws = nullptr;
    rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    struct Rel { DeviceCtx* d; Workspace* w; ~Rel() { ws_release(d, w); } } rel{dc, ws};
    cudaStream_t st = ws->st;
    const size_t n0 = m->n, n1 = n0 + k;
    const size_t N1 = (n1 + TB - 1) / TB * TB;
    if (N1 > m->cap) {
        size_t want = std::max(N1, (m->cap * 5 / 4 + TB - 1) / TB * TB);
        rc = grow_capacity(m, want, st);
        if (rc) return rc;
    }
    ModelDev& md = m->devs[0];
    const size_t ld = m->cap;
    const size_t need = append_workspace_doubles(ld);
    if (m->aws_dbl < need) {
        cudaFree(m->aws); m->aws = nullptr; m->aws_dbl = 0;
        CU(cudaMalloc((void**)&m->aws, need * sizeof(double)));
        m->aws_dbl = need;
    }
    CU(cudaEventRecord(ws->ev[0], st));
    CU(cudaMemcpyAsync(md.xyz + n0, x, k * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(md.xyz + ld + n0, y, k * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(md.xyz + 2 * ld + n0, z, k * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->label + n0, label, k * sizeof(double), cudaMemcpyHostToDevice, st));
    if (sigma2) CU(cudaMemcpyAsync(m->s2 + n0, sigma2, k * sizeof(double), cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(m->s2 + n0, 0, k * sizeof(double), st));
    CU(cudaEventRecord(ws->ev[1], st));
    for (size_t o = 0; o < k; o += 32) {
        const int kk = (int)std::min<size_t>(32, k - o);
        CU(launch_append_slab(md.xyz, ld, m->s2, (int)(n0 + o), kk, m->L, md.linv, m->Dinv, m->aws, ld, m->kp, o == 0, st));
    }
    // No host round trip between the slabs and the solve: alpha is solved into a scratch vector from whatever the slabs
    // left behind and committed on the device only if the sticky flag is clear, so that a failed append leaves alpha
    // (like L, L^-1 and Dinv) untouched; the flag is read once, at the end.
    const int nb1 = (int)(N1 / TB);
    // alpha = X^T (X y) through the inverse factor that the append has just brought up to date (two bandwidth-bound
    // passes, no dependency chain), then the same refinement as after a fit with the correction also through X.
    // GPR_APPEND_TRSV=1 uses the triangular solves over L instead.
    static const bool use_trsv = getenv("GPR_APPEND_TRSV") && atoi(getenv("GPR_APPEND_TRSV")) != 0;
    const size_t sol_dbl = (size_t)(32 + 1) * m->cap;                  // sized by the capacity: no reallocation per call
    rc = ws_reserve(&ws->mpart, &ws->mpart_dbl, residual_scratch_doubles((int)m->cap) + 3 * m->cap + sol_dbl);
    if (rc) return rc;
    double* rr = ws->mpart + residual_scratch_doubles((int)N1);
    double* dd = rr + N1;
    double* anew = dd + N1;
    double* sol = anew + N1;
    const int* dflag = append_flag_ptr(m->aws, ld);
    CU(cudaMemsetAsync(anew, 0, N1 * sizeof(double), st));
    if (use_trsv) {
        CU(launch_trsv(0, m->L, ld, nb1, m->Dinv, m->label, m->zfwd, m->scratch, dc->num_sms, st));
        CU(launch_trsv(1, m->L, ld, nb1, m->Dinv, m->zfwd, anew, m->scratch, dc->num_sms, st));
    } else {
        CU(launch_solve_with_inverse(md.linv, ld, (int)n1, m->label, anew, sol, st));
    }
    for (int it = 0; it < std::max(1, ctx->refine_steps); ++it) {      // at least one step: it also absorbs the explicit-inverse rounding
        CU(launch_residual(md.xyz, ld, m->s2, m->label, anew, (int)n1, (int)N1, ws->mpart, rr, m->kp, st));
        if (use_trsv) {
            CU(launch_trsv(0, m->L, ld, nb1, m->Dinv, rr, m->zfwd, m->scratch, dc->num_sms, st));
            CU(launch_trsv(1, m->L, ld, nb1, m->Dinv, m->zfwd, dd, m->scratch, dc->num_sms, st));
        } else {
            CU(launch_solve_with_inverse(md.linv, ld, (int)n1, rr, dd, sol, st));
        }
        CU(launch_axpy1(anew, dd, (int)n1, st));
    }
    CU(launch_commit_if_clear(dflag, anew, md.alpha, (int)n1, st));
    CU(cudaEventRecord(ws->ev[2], st));
    std::vector<double> new_alpha(n1);
    CU(cudaMemcpyAsync(new_alpha.data(), md.alpha, n1 * sizeof(double), cudaMemcpyDeviceToHost, st));
    int flag = 0;
    int flags[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (use_trsv) CU(cudaMemcpyAsync(flags, m->scratch, sizeof(flags), cudaMemcpyDeviceToHost, st));   // the control words belong to the last flag-chained kernel
    CU(cudaStreamSynchronize(st));
    if (flag != 0) {
        // roll back: the rows of the slabs that were committed before the failing one become padding again
        CU(launch_identity_rows(m->L, md.linv, ld, (int)n0, (int)n1, (int)N1, st));
        CU(launch_dinv_from_x(md.linv, ld, (int)(n0 / TB), (int)(N1 / TB - n0 / TB), m->Dinv, st));
        for (int c = 0; c < 3; ++c) CU(cudaMemsetAsync(md.xyz + c * ld + n0, 0, k * sizeof(double), st));
        CU(cudaMemsetAsync(m->label + n0, 0, k * sizeof(double), st));
        CU(cudaMemsetAsync(m->s2 + n0, 0, k * sizeof(double), st));
        CU(cudaStreamSynchronize(st));
        g_pivot = flag;
        char b[256];
        snprintf(b, sizeof b, "covariance matrix is not positive definite after the append: pivot %d of %zu is <= 0", flag, n1);
        return fail(GPR_ERR_NOT_SPD, b);
    }
    if (flags[2] != 0) return fail(GPR_ERR_CUDA, "triangular solve kernel aborted (dependency wait timed out)");
    m->h_alpha.swap(new_alpha);
    host_append(m, x, y, z, label, sigma2, k);
    m->n = n1; m->n_spd = n1; m->N = N1; m->nb = nb1;
    for (size_t di = 1; di < m->devs.size(); ++di) { m->devs[di].have = false; m->devs[di].have_linv = false; m->devs[di].have_fac = false; }
    std::lock_guard<std::mutex> lk(ctx->tmu);
    ctx->timings.h2d_ms = ev_ms(ws->ev[0], ws->ev[1]);
    ctx->timings.append_ms = ev_ms(ws->ev[1], ws->ev[2]);
    return GPR_OK;
}

}  // extern "C"

// Incremental update of a model with an indefinite tail block (the node's real setting: cb_update adds touch points to a
// ThinPlate(2.0) model, src/gp_node.cpp:652-763, and refits from scratch).  The new points join the positive definite
// leading block — rows of L and of X = L^-1 are appended by the same slab kernels as for an SPD model, with the tail
// points moved behind them in the internal order — and the trailing block is eliminated again against the extended X
// (B, Z, S^-1, alpha: bandwidth-bound passes over X, gpr_tail.cu), instead of two n^3/3 factorisations.
// Returns 1000 (no error set) when the new points do not fit the leading block (non-positive pivot): the caller refits,
// which re-runs the eviction logic.  Caller holds m->mu.
constexpr int TAIL_APPEND_FALLBACK = 1000;
static int append_tail_incremental(gpr_model* m, const double* x, const double* y, const double* z, const double* label,
                                   const double* sigma2, size_t k) {
    gpr_ctx* ctx = m->ctx;
    DeviceCtx* dc = ctx->devs[0];
    CU(cudaSetDevice(dc->dev));
    ModelDev& md0 = m->devs[0];
    if (!md0.have_linv || !md0.have_tail) return TAIL_APPEND_FALLBACK;
    Workspace* ws = nullptr;
    int rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    struct Rel { DeviceCtx* d; Workspace* w; ~Rel() { ws_release(d, w); } } rel{dc, ws};
    cudaStream_t st = ws->st;
    const size_t p = m->n_spd, mt = m->n_tail, n0 = m->n, n1 = n0 + k;
    const size_t N1 = (n1 + TB - 1) / TB * TB;
    if (N1 > m->cap) {
        rc = grow_capacity(m, std::max(N1, (m->cap * 5 / 4 + TB - 1) / TB * TB), st);
        if (rc) return rc;
    }
    ModelDev& md = m->devs[0];
    const size_t ld = m->cap;
    const size_t need = append_workspace_doubles(ld);
    if (m->aws_dbl < need) {
        cudaFree(m->aws); m->aws = nullptr; m->aws_dbl = 0;
        CU(cudaMalloc((void**)&m->aws, need * sizeof(double)));
        m->aws_dbl = need;
    }
    CU(cudaEventRecord(ws->ev[0], st));
    // everything behind the leading block becomes padding in L / X (tile rows beyond the leading block may hold what a
    // failed first factorisation attempt left there), with matching diagonal-block inverses
    CU(launch_identity_rows(m->L, md.linv, ld, (int)p, (int)ld, (int)ld, st));
    CU(launch_dinv_from_x(md.linv, ld, (int)(p / TB), (int)(ld / TB - p / TB), m->Dinv, st));
    // internal order: [leading block | new points | tail]: move the tail's coordinates, labels and noise k places up
    rc = ws_reserve(&ws->io, &ws->io_cap, std::max((size_t)14 * TB, 5 * mt));
    if (rc) return rc;
    double* tmp = ws->io;
    const double* srcs[5] = {md.xyz + p, md.xyz + ld + p, md.xyz + 2 * ld + p, m->label + p, m->s2 + p};
    double* dsts[5] = {md.xyz + p + k, md.xyz + ld + p + k, md.xyz + 2 * ld + p + k, m->label + p + k, m->s2 + p + k};
    for (int c = 0; c < 5; ++c) CU(cudaMemcpyAsync(tmp + c * mt, srcs[c], mt * sizeof(double), cudaMemcpyDeviceToDevice, st));
    for (int c = 0; c < 5; ++c) CU(cudaMemcpyAsync(dsts[c], tmp + c * mt, mt * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(md.xyz + p, x, k * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(md.xyz + ld + p, y, k * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(md.xyz + 2 * ld + p, z, k * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->label + p, label, k * sizeof(double), cudaMemcpyHostToDevice, st));
    if (sigma2) CU(cudaMemcpyAsync(m->s2 + p, sigma2, k * sizeof(double), cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(m->s2 + p, 0, k * sizeof(double), st));
    CU(cudaEventRecord(ws->ev[1], st));
    for (size_t o = 0; o < k; o += 32) {
        const int kk = (int)std::min<size_t>(32, k - o);
        CU(launch_append_slab(md.xyz, ld, m->s2, (int)(p + o), kk, m->L, md.linv, m->Dinv, m->aws, ld, m->kp, o == 0, st));
    }
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, append_flag_ptr(m->aws, ld), sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (flag != 0) return TAIL_APPEND_FALLBACK;            // the device state is rebuilt from the host copies by the refit
    // bookkeeping of the new shape, then the trailing block against the extended factor
    std::vector<size_t> np;
    np.reserve(n1);
    for (size_t i = 0; i < p; ++i) np.push_back(m->perm.empty() ? i : m->perm[i]);
    for (size_t i = 0; i < k; ++i) np.push_back(n0 + i);
    for (size_t i = p; i < n0; ++i) np.push_back(m->perm.empty() ? i : m->perm[i]);
    m->n_spd = p + k; m->nb = (int)((p + k + TB - 1) / TB); m->n = n1; m->N = N1;
    rc = tail_eliminate(m, dc, ws, false);
    if (rc) {                                              // e.g. a singular Schur complement: refit decides
        m->n_spd = p; m->nb = (int)((p + TB - 1) / TB); m->n = n0; m->N = (n0 + TB - 1) / TB * TB;
        return rc == GPR_ERR_NOT_SPD ? TAIL_APPEND_FALLBACK : rc;
    }
    CU(cudaEventRecord(ws->ev[2], st));
    m->perm.swap(np);
    host_append(m, x, y, z, label, sigma2, k);
    std::vector<double> a_int(n1);
    CU(cudaMemcpyAsync(a_int.data(), md.alpha, n1 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    m->h_alpha.assign(n1, 0.0);
    for (size_t i = 0; i < n1; ++i) m->h_alpha[m->perm[i]] = a_int[i];
    for (size_t di = 1; di < m->devs.size(); ++di) {       // copies on the other devices: re-made on demand (the tail slabs' leading dimension may have changed)
        ModelDev& d = m->devs[di];
        cudaSetDevice(d.dev);
        cudaFree(d.tZ); cudaFree(d.tSinv);
        d.tZ = d.tSinv = nullptr;
        d.have = d.have_linv = d.have_tail = d.have_fac = false;
    }
    cudaSetDevice(dc->dev);
    std::lock_guard<std::mutex> lk(ctx->tmu);
    ctx->timings.h2d_ms = ev_ms(ws->ev[0], ws->ev[1]);
    ctx->timings.append_ms = ev_ms(ws->ev[1], ws->ev[2]);
    return GPR_OK;
}

extern "C" {

int gpr_append(gpr_ctx* ctx, gpr_model* m, const double* x, const double* y, const double* z, const double* label,
               const double* sigma2, size_t k) {
    if (!ctx) return fail(GPR_ERR_INVALID, "null context");
    if (!m) return fail(GPR_ERR_INVALID, "Empty model pointer");
    if (!x || !y || !z || !label || k == 0) return fail(GPR_ERR_INVALID, "All input data is empty!");
    if (m->replica) return fail(GPR_ERR_INVALID, "cannot append to a replica model");
    std::lock_guard<std::mutex> lk(m->mu);
    const char* force = getenv("GPR_APPEND_REFIT");
    // a permuted internal order (points were evicted during the fit, with or without a tail block left) refits: the
    // incremental path works in the caller's order
    const bool forced = force && atoi(force) != 0;
    const bool small = k <= 256 && 8 * k <= m->n;
    if (!forced && small && m->n_tail > 0 && m->n_spd >= 8 * k) {
        // indefinite-tail model: rows appended to the leading block, trailing block eliminated again
        const int rc = append_tail_incremental(m, x, y, z, label, sigma2, k);
        if (rc != TAIL_APPEND_FALLBACK) return rc;
    }
    const bool refit = forced || !small || m->n_tail > 0 || !m->perm.empty();
    if (!refit) return append_incremental(m, x, y, z, label, sigma2, k);
    // Large batches: append on the host and refit, like the reference (gp_regressor.hpp:442-459).
    const size_t p = m->hx.size();
    const bool had_s2 = m->has_s2;
    host_append(m, x, y, z, label, sigma2, k);
    const bool normals = m->with_normals;
    m->with_normals = false;                 // :462-477: normals are not refreshed by update()
    std::vector<double> keep = m->h_normals;
    const size_t keep_n = m->n_normals;
    int rc = fit_from_host(m, true);
    { std::lock_guard<std::mutex> tl(ctx->tmu); ctx->timings.append_ms = 0.0; }   // no incremental update ran
    m->with_normals = normals;
    m->h_normals = keep; m->n_normals = keep_n;
    if (rc) {
        // leave the model as it was: drop the appended points and refit the old set
        const int code = rc; const std::string msg = g_err; const long long piv = g_pivot;
        m->hx.resize(p); m->hy.resize(p); m->hz.resize(p); m->hlabel.resize(p);
        if (had_s2) m->hs2.resize(p); else { m->hs2.clear(); m->has_s2 = false; }
        m->with_normals = false;
        fit_from_host(m, true);
        m->with_normals = normals; m->h_normals = keep; m->n_normals = keep_n;
        g_err = msg; g_pivot = piv;
        return code;
    }
    return rc;
}

int gpr_model_reserve(gpr_ctx* ctx, gpr_model* m, size_t capacity) {
    if (!ctx || !m) return fail(GPR_ERR_INVALID, "Empty model pointer");
    if (m->replica) return fail(GPR_ERR_INVALID, "cannot reserve on a replica model");
    std::lock_guard<std::mutex> lk(m->mu);
    const size_t want = (capacity + TB - 1) / TB * TB;
    if (want <= m->cap) return GPR_OK;
    if (m->n_tail > 0) return GPR_OK;        // updates of a model with an indefinite tail block refit: nothing to reserve
    DeviceCtx* dc = ctx->devs[0];
    CU(cudaSetDevice(dc->dev));
    Workspace* ws = nullptr;
    int rc = ws_acquire(dc, &ws);
    if (rc) return rc;
    rc = grow_capacity(m, want, ws->st);
    ws_release(dc, ws);
    return rc;
}

int gpr_model_state_get(gpr_ctx* ctx, gpr_model* m, int with_linv, gpr_model_state* out) {
    if (!ctx || !m || !out) return fail(GPR_ERR_INVALID, "null pointer");
    if (with_linv & 1) { int rc = ensure_on_device(m, 0, true); if (rc) return rc; }
    out->n = m->n; out->padded_n = m->N; out->ld = m->cap; out->kernel = m->kernel; out->R = m->R;
    out->xyz = m->devs[0].xyz; out->alpha = m->devs[0].alpha;
    out->linv = m->devs[0].have_linv ? m->devs[0].linv : nullptr;
    out->n_tail = m->n_tail; out->tail_pad = (size_t)m->mp;
    out->tail_z = m->n_tail ? m->devs[0].tZ : nullptr;
    out->tail_sinv = m->n_tail ? m->devs[0].tSinv : nullptr;
    out->lfac = m->devs[0].have_fac ? m->devs[0].lfac : nullptr;
    out->dinv = out->lfac ? m->devs[0].dinv : nullptr;
    return GPR_OK;
}

int gpr_model_create_replica_tail(gpr_ctx* ctx, size_t n, size_t n_tail, gpr_kernel_t kernel, double R, int with_linv,
                                  gpr_model** out);

int gpr_model_create_replica(gpr_ctx* ctx, size_t n, gpr_kernel_t kernel, double R, int with_linv, gpr_model** out) {
    return gpr_model_create_replica_tail(ctx, n, 0, kernel, R, with_linv, out);
}

int gpr_model_create_replica_tail(gpr_ctx* ctx, size_t n, size_t n_tail, gpr_kernel_t kernel, double R, int with_linv,
                                  gpr_model** out) {
    if (!ctx || !out || n == 0 || n_tail >= n || n_tail > MAX_TAIL) return fail(GPR_ERR_INVALID, "null pointer, empty model or bad tail size");
    gpr_model* m = new gpr_model();
    m->ctx = ctx; m->kernel = kernel; m->kp = make_kp(kernel); m->k0 = kernel_at_zero(m->kp); m->R = R;
    m->replica = true;
    m->n = n; m->n_spd = n; m->N = (n + TB - 1) / TB * TB; m->nb = (int)(m->N / TB); m->cap = m->N;
    m->devs.assign(ctx->devs.size(), ModelDev());
    for (size_t i = 0; i < ctx->devs.size(); ++i) m->devs[i].dev = ctx->devs[i]->dev;
    ModelDev& md = m->devs[0];
    auto bail = [&](int rc) { free_factor(m); delete m; return rc; };
    if (cudaSetDevice(md.dev) != cudaSuccess) return bail(fail(GPR_ERR_CUDA, "cudaSetDevice failed"));
    if (cudaMalloc((void**)&md.xyz, 3 * m->N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&md.alpha, m->N * sizeof(double)) != cudaSuccess)
        return bail(fail(GPR_ERR_OOM, "out of device memory"));
    md.have = true;
    if (with_linv & 1) {
        if (cudaMalloc((void**)&md.linv, m->N * m->N * sizeof(double)) != cudaSuccess)
            return bail(fail(GPR_ERR_OOM, "out of device memory"));
        md.have_linv = true;
    }
    if (with_linv & 2) {
        if (n_tail > 0) return bail(fail(GPR_ERR_INVALID, "a replica of an indefinite-tail model takes L^-1, not the factor"));
        if (cudaMalloc((void**)&md.lfac, m->N * m->N * sizeof(double)) != cudaSuccess ||
            cudaMalloc((void**)&md.dinv, (size_t)m->nb * TB * TB * sizeof(double)) != cudaSuccess)
            return bail(fail(GPR_ERR_OOM, "out of device memory"));
        md.own_fac = true; md.have_fac = true;
    }
    if (n_tail > 0) {
        // the trailing pivot block of an indefinite matrix (gpr_tail.cu): L / L^-1 describe the leading n - n_tail points
        m->n_spd = n - n_tail; m->n_tail = n_tail; m->nb = (int)((m->n_spd + TB - 1) / TB);
        m->mp = 32 * (int)((n_tail + 31) / 32);
        if (cudaMalloc((void**)&md.tZ, (size_t)(m->mp / 32) * m->cap * 32 * sizeof(double)) != cudaSuccess ||
            cudaMalloc((void**)&md.tSinv, (size_t)m->mp * m->mp * sizeof(double)) != cudaSuccess)
            return bail(fail(GPR_ERR_OOM, "out of device memory"));
        md.have_tail = true;
    }
    *out = m;
    return GPR_OK;
}

// ---- self-tests ---------------------------------------------------------------------------------
int gpr_selftest_gemm(const double* hA, const double* hB, int b_kmajor, double* hC, int mt, int nt, int k) {
    if (k % (2 * KT)) return fail(GPR_ERR_INVALID, "k must be a multiple of 32 (the mainloop works on pairs of k16 stages)");
    const size_t M = (size_t)mt * TB, Nn = (size_t)nt * TB;
    double *A, *B, *C;
    CU(cudaMalloc((void**)&A, M * k * sizeof(double)));
    CU(cudaMalloc((void**)&B, Nn * k * sizeof(double)));
    CU(cudaMalloc((void**)&C, M * Nn * sizeof(double)));
    CU(cudaMemcpy(A, hA, M * k * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(B, hB, Nn * k * sizeof(double), cudaMemcpyHostToDevice));
    CU(launch_gemm_selftest(A, M, B, b_kmajor ? (size_t)k : Nn, b_kmajor, C, M, mt, nt, k, 0));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(hC, C, M * Nn * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(A); cudaFree(B); cudaFree(C);
    return GPR_OK;
}

int gpr_selftest_leaf(double* h_tile, double* h_inv, int* info) {
    double *T, *I; int* d_info;
    CU(cudaMalloc((void**)&T, TB * TB * sizeof(double)));
    CU(cudaMalloc((void**)&I, TB * TB * sizeof(double)));
    CU(cudaMalloc((void**)&d_info, sizeof(int)));
    CU(cudaMemcpy(T, h_tile, TB * TB * sizeof(double), cudaMemcpyHostToDevice));
    CU(launch_leaf_selftest(T, I, d_info, 0));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h_tile, T, TB * TB * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_inv, I, TB * TB * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(info, d_info, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(T); cudaFree(I); cudaFree(d_info);
    return GPR_OK;
}

int gpr_selftest_factor(double* hA, int nb, double* h_linv, int serial, long long* pivot) {
    const size_t N = (size_t)nb * TB;
    double *A, *D, *X = nullptr; int* scratch;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    CU(cudaMalloc((void**)&A, N * N * sizeof(double)));
    CU(cudaMalloc((void**)&D, (size_t)nb * TB * TB * sizeof(double)));
    CU(cudaMalloc((void**)&scratch, (8 + (size_t)nb * nb) * sizeof(int)));
    CU(cudaMemcpy(A, hA, N * N * sizeof(double), cudaMemcpyHostToDevice));
    if (serial >= 100) {
        // INT8-assisted factorisation: serial = 100 * (tile columns per panel) + (digit slices)
        const int P = serial / 100, S = serial % 100;
        signed char* Ls; int* ctrl;
        CU(cudaMalloc((void**)&Ls, ozaki_fit_workspace_bytes(S, N)));
        CU(cudaMalloc((void**)&ctrl, 4 * sizeof(int)));
        CU(launch_cholesky_int8(A, N, nb, D, scratch, sms, 0, nullptr, Ls, S, P, 0, ctrl));
        CU(cudaDeviceSynchronize());
        int hc[2];
        CU(cudaMemcpy(hc, ctrl, sizeof hc, cudaMemcpyDeviceToHost));
        cudaFree(Ls); cudaFree(ctrl);
        if (hc[1]) return fail(GPR_ERR_CUDA, "int8 update kernel timed out");
    } else {
        CU(launch_cholesky(A, N, nb, D, scratch, sms, serial, 0));
    }
    CU(cudaDeviceSynchronize());
    int info[4];
    CU(cudaMemcpy(info, scratch, sizeof info, cudaMemcpyDeviceToHost));
    if (pivot) *pivot = info[1];
    if (info[2] && !info[1]) return fail(GPR_ERR_CUDA, "cholesky kernel aborted");
    CU(cudaMemcpy(hA, A, N * N * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_linv && !info[1]) {
        CU(cudaMalloc((void**)&X, N * N * sizeof(double)));
        CU(cudaMemset(X, 0, N * N * sizeof(double)));
        CU(launch_linv(A, X, N, nb, D, scratch, sms, 0));
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(info, scratch, sizeof info, cudaMemcpyDeviceToHost));
        if (info[2]) return fail(GPR_ERR_CUDA, "L^-1 kernel aborted");
        CU(cudaMemcpy(h_linv, X, N * N * sizeof(double), cudaMemcpyDeviceToHost));
        cudaFree(X);
    }
    cudaFree(A); cudaFree(D); cudaFree(scratch);
    return info[1] ? GPR_ERR_NOT_SPD : GPR_OK;
}

// Timeline of the tile-task Cholesky on an SPD test matrix of n_tiles x n_tiles tiles: 4 globaltimer
// stamps (ns) per task in task order (claimed, accumulation done, tile solved, flag published).
int gpr_selftest_factor_trace(int nb, long long* h_trace, long long* leaf_cycles) {
    const size_t N = (size_t)nb * TB;
    std::vector<double> hA(N * N, 0.0);
    for (size_t c = 0; c < N; ++c)
        for (size_t r = c; r < N; ++r) hA[c * N + r] = (r == c) ? 4.0 + 0.001 * (double)(r % 7) : 1.0 / (1.0 + (double)(r - c));
    double *A, *D; int* scratch; long long* tr;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t ntasks = (size_t)nb * (nb + 1) / 2;
    CU(cudaMalloc((void**)&A, N * N * sizeof(double)));
    CU(cudaMalloc((void**)&D, (size_t)nb * TB * TB * sizeof(double)));
    CU(cudaMalloc((void**)&scratch, (8 + (size_t)nb * nb) * sizeof(int)));
    CU(cudaMalloc((void**)&tr, 4 * ntasks * sizeof(long long)));
    for (int rep = 0; rep < 2; ++rep) {
        CU(cudaMemcpy(A, hA.data(), N * N * sizeof(double), cudaMemcpyHostToDevice));
        CU(cudaMemset(tr, 0, 4 * ntasks * sizeof(long long)));
        CU(launch_cholesky(A, N, nb, D, scratch, sms, 0, 0, tr));
        CU(cudaDeviceSynchronize());
    }
    int info[4];
    CU(cudaMemcpy(info, scratch, sizeof info, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_trace, tr, 4 * ntasks * sizeof(long long), cudaMemcpyDeviceToHost));
    if (leaf_cycles) {
        double* I; int* di; long long* cy;
        CU(cudaMalloc((void**)&I, TB * TB * sizeof(double)));
        CU(cudaMalloc((void**)&di, sizeof(int)));
        CU(cudaMalloc((void**)&cy, 8 * sizeof(long long)));
        std::vector<double> t(TB * TB);
        for (int c = 0; c < TB; ++c) for (int r = 0; r < TB; ++r) t[c * TB + r] = (r == c) ? 4.0 : 1.0 / (1.0 + std::abs(r - c));
        for (int rep = 0; rep < 2; ++rep) {
            CU(cudaMemcpy(A, t.data(), TB * TB * sizeof(double), cudaMemcpyHostToDevice));
            CU(launch_leaf_selftest(A, I, di, 0, cy));
            CU(cudaDeviceSynchronize());
        }
        CU(cudaMemcpy(leaf_cycles, cy, 8 * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(I); cudaFree(di); cudaFree(cy);
    }
    cudaFree(A); cudaFree(D); cudaFree(scratch); cudaFree(tr);
    return info[1] || info[2] ? GPR_ERR_CUDA : GPR_OK;
}

// INT8 tensor-core engine self-test (gpr_ozaki.cu): raw level accumulators C[l] = sum_{t+u=l} A_t B_u^T for int8 slice
// tensors A [S][M][K], B [S][N][K] (host, K contiguous); M multiple of 128, N of 64, K of 64; tri: A lower triangular by
// 128-row tiles (row tile r only visits k < 128 (r + 1)).  hC: [levels][M][N] int32.
int gpr_selftest_i8gemm(const signed char* hA, const signed char* hB, int S, int levels, int M, int Nq, int K, int tri, int skip_zero_blocks,
                        int* hC) {
    if (!hA || !hB || !hC || M % 128 || Nq % 16 || K % 64 || S < 1 || S > 8 || levels != S) return fail(GPR_ERR_INVALID, "bad shape");
    signed char *A, *B; int *C, *ctrl; double *scale, *partial;
    const size_t qpad = (size_t)(Nq + 127) / 128 * 128;
    CU(cudaMalloc((void**)&A, (size_t)S * M * K));
    CU(cudaMalloc((void**)&B, (size_t)S * Nq * K));
    CU(cudaMalloc((void**)&C, (size_t)S * M * Nq * sizeof(int)));
    CU(cudaMalloc((void**)&ctrl, 4 * sizeof(int)));
    CU(cudaMalloc((void**)&scale, (size_t)M * sizeof(double)));
    CU(cudaMalloc((void**)&partial, (size_t)(M / 128) * qpad * sizeof(double)));
    CU(cudaMemcpy(A, hA, (size_t)S * M * K, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(B, hB, (size_t)S * Nq * K, cudaMemcpyHostToDevice));
    CU(cudaMemset(C, 0, (size_t)S * M * Nq * sizeof(int)));
    std::vector<double> ones(M, 1.0);
    CU(cudaMemcpy(scale, ones.data(), (size_t)M * sizeof(double), cudaMemcpyHostToDevice));
    unsigned char* nz = nullptr;
    if (skip_zero_blocks) {
        CU(cudaMalloc((void**)&nz, (size_t)(M / 128) * (K / 64)));
        CU(launch_ozaki_mask(A, (size_t)K, (size_t)M * K, S, M / 128, K / 64, nz, (size_t)(K / 64), 0));
    }
    // the tensor of B has exactly Nq rows: query tiles reaching beyond it are zero-filled by TMA
    CU(launch_ozaki_product(A, (size_t)K, (size_t)M * K, M / 128, B, (size_t)K, (size_t)Nq * K, (size_t)Nq, Nq, qpad, (size_t)K, tri, S, 0, scale,
                            1.0, partial, ctrl, C, (size_t)Nq, 0, nz, (size_t)(K / 64)));
    CU(cudaDeviceSynchronize());
    int hctrl[2];
    CU(cudaMemcpy(hctrl, ctrl, sizeof hctrl, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hC, C, (size_t)S * M * Nq * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(ctrl); cudaFree(scale); cudaFree(partial); cudaFree(nz);
    if (hctrl[1] != 0) return fail(GPR_ERR_CUDA, "INT8 tensor-core kernel aborted (a pipeline wait timed out)");
    return GPR_OK;
}

int gpr_selftest_peak(int which, int ctas_per_sm, double* tflops) {
    if (!tflops) return fail(GPR_ERR_INVALID, "null pointer");
    CU(run_peak_probe(which, ctas_per_sm, tflops));
    return GPR_OK;
}

}  // extern "C"

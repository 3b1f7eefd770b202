// gpr_factor.cu — K2: blocked FP64 Cholesky and the triangular inverse L^-1, each as ONE persistent
// tile-task kernel on the FP64 tensor pipe (DMMA).
//
// Replaces Eigen's unblocked, pivoted, single-threaded LDLT::compute in the reference
// (/root/reference/include/gp_regression/gp_regressor.hpp:161-162, :457-458).  On the SPD inputs of
// every BASELINE config LDLT and LLT agree to cond(K)*eps; a non-positive pivot is
// reported (never NaN), see SURVEY F2.
//
// Algorithm (left-looking by 128x128 tiles, dependency flags in global memory):
//   task (i,j), i >= j, claimed from an atomic counter in column order (diagonal tasks one column ahead,
//   see chol_task_decode):
//     acc  = sum_{k<j} L_ik L_jk^T                 DMMA mainloop; waits on ready[i][k], ready[j][k]
//     T    = K_ij - acc                            -> shared memory
//     i==j : L_jj = chol(T) in shared memory (warp-level 16x16 diagonal blocks, DMMA trailing
//            updates), Dinv_j = L_jj^-1, both written to global; ready[j][j] = 1
//     i>j  : wait ready[j][j]; L_ij = T * Dinv_j^T  (DMMA, T resident in shared memory); ready[i][j] = 1
//   Every tile is produced by one CTA in a fixed k order, so the factor is bit-reproducible.
//   Every dependency of a task has a smaller task index, hence is held by a CTA that is already
//   running: the flag waits cannot deadlock whatever the number of resident CTAs.
#include "gpr_mma.cuh"
#include "gpr_leaf.cuh"
#include "gpr_kernels.h"

namespace gpr {

// ---------------------------------------------------------------------------------------------
// Cholesky tile-task kernel.
// ---------------------------------------------------------------------------------------------
struct CholArgs {
    double* A;        // N x N column-major, lower tiles hold K on entry and L on exit
    size_t ld;
    int nb;           // N / 128
    double* Dinv;     // nb tiles of 128x128: inverse of each diagonal block of L
    int* ready;       // flags, [i*ready_ld + j]
    int ready_ld;     // nb for a whole factorisation; the enclosing matrix's nb for a panel launch
    int ncols;        // only the first ncols tile columns are factorised (nb for a whole factorisation): panel launches of the
                      // INT8-assisted factorisation (launch_cholesky_int8), where A, Dinv and ready point at the panel's corner
    int row0;         // global row of A's first row (pivot reports)
    int* counter;     // tasks claimed so far by this launch
    int task_begin;   // tasks [task_begin, task_end) are executed by this launch
    int task_end;
    int* info;        // 0, or 1 + global index of the first non-positive pivot
    int* abort;       // raised on failure so that waiting CTAs leave
    long long* trace; // optional: 4 globaltimer stamps per task (claim, accumulated, solved, published)
    // Replication fused into the producer (multi-GPU predict): every finished tile of L and of Dinv is also
    // stored into the same position of up to MAX_CHOL_PEERS peer buffers (other GPUs' replicas, mapped through CUDA IPC /
    // peer access, same leading dimension), so the replicas are complete when the factorisation ends — the n x n
    // broadcast that used to follow the fit disappears from the critical path.  Posted NVLink writes, no handshake:
    // the consumers synchronise with the end of the kernel (stream sync + the launcher's barrier).
    int n_peers;
    double* peerA[MAX_CHOL_PEERS];
    double* peerDinv[MAX_CHOL_PEERS];
};

// Task order.  Column j contributes, in this order, (j+1,j), (j+1,j+1), (j+2,j), ..., (nb-1,j); task 0 is
// (0,0).  The diagonal task of column j+1 is therefore claimed one whole column ahead of the tasks
// that need its result (L_{j+1,j+1}^-1): its in-CTA Cholesky + inverse, the only serial piece of the
// algorithm, overlaps the accumulation of the column-j tasks instead of stalling every task of column
// j+1 behind it.  Dependencies of (i,j): (i,k),(j,k) for k<j (earlier columns), (j,j) (second task of
// column j-1), and for a diagonal task (j,j-1), which is the task just before it — all smaller indices.
__host__ __device__ __forceinline__ void chol_task_decode(int task, int nb, int& i, int& j) {
    if (task == 0) { i = 0; j = 0; return; }
    task -= 1;
    j = 0;
    while (task >= nb - j) { task -= nb - j; ++j; }
    if (task == 0) { i = j + 1; }
    else if (task == 1) { i = j + 1; j = j + 1; }
    else { i = j + task; }
}

__global__ void __launch_bounds__(NTHREADS, 1) chol_tiles_kernel(CholArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_task, s_abort, s_fail, s_upto, s_i, s_j;
    __shared__ double s_inv[TB];
    const int tid = threadIdx.x;
    // The tile coordinates live in shared memory and are re-read (volatile) by each phase of a task, so that
    // nothing but the accumulators, the fragments and the copy addresses is live across the DMMA loop: with
    // them held in registers the loop ran out of registers, ptxas rotated the accumulators and the address
    // updates of the asynchronous copies stalled on the previous copies (long-scoreboard stalls, ~8 %).
#define GPR_SH(x) (*(volatile int*)&(x))
    for (;;) {
        if (tid == 0) {
            const int t = a.task_begin + atomicAdd(a.counter, 1);
            int ti = 0, tj = 0;
            if (t < a.task_end) chol_task_decode(t, a.nb, ti, tj);
            if (tj >= a.ncols) ti = -1;                      // diagonal task of the column after a panel: not part of this launch
            s_task = t; s_i = ti; s_j = tj;
            s_abort = ld_volatile(a.abort) != 0;
            s_fail = 1 << 20;
            s_upto = 1 << 30;
            if (a.trace && t < a.task_end) a.trace[4 * (size_t)t + 0] = globaltimer_ns();
        }
        __syncthreads();
        if (GPR_SH(s_task) >= a.task_end || s_abort) return;
        if (GPR_SH(s_i) < 0) { __syncthreads(); continue; }

        Acc acc;
        acc_zero(acc);
        {
            // One bulk look at the readiness flags of both row panels: the k-blocks [0, upto) are complete
            // (almost always all but the last one or two).  They are accumulated by the wait-free mainloop
            // (the same instantiation as the variance kernel).
            const int i = GPR_SH(s_i), j = GPR_SH(s_j);
            const int* fi = a.ready + (size_t)i * a.ready_ld;
            const int* fj = a.ready + (size_t)j * a.ready_ld;
            const int upto = ready_prefix(j, [&](int t) { return ld_acquire(fi + t) != 0 && ld_acquire(fj + t) != 0; }, &s_upto);
            if (!tile_mainloop<STREAM_M, STREAM_M>(acc, a.A + (size_t)i * TB, a.ld, a.A + (size_t)j * TB, a.ld, 8 * upto, smem,
                                                   &s_abort, NoWait())) return;
        }
        // the few blocks beyond the prefix are awaited and accumulated one 128-wide block at a time
        for (int kb = GPR_SH(s_upto); kb < GPR_SH(s_j); ++kb) {
            const int i = GPR_SH(s_i), j = GPR_SH(s_j);
            if (tid == 0) {
                if (!spin_wait(a.ready + (size_t)i * a.ready_ld + kb, a.abort) ||
                    (i != j && !spin_wait(a.ready + (size_t)j * a.ready_ld + kb, a.abort))) s_abort = 1;
            }
            __syncthreads();
            if (s_abort) return;
            const size_t off = (size_t)kb * TB * a.ld;
            if (!tile_mainloop<STREAM_M, STREAM_M>(acc, a.A + (size_t)i * TB + off, a.ld, a.A + (size_t)j * TB + off, a.ld, 8,
                                                   smem, &s_abort, NoWait())) return;
        }

        const int i = GPR_SH(s_i), j = GPR_SH(s_j), task_id = GPR_SH(s_task);
        const TileCoord tc;
        if (a.trace && tid == 0) a.trace[4 * (size_t)task_id + 1] = globaltimer_ns();
        double* T = smem;   // region 0, column-major pitch PM
        double* Gij = a.A + (size_t)j * TB * a.ld + (size_t)i * TB;
        residual_to_smem<false>(acc, Gij, a.ld, T, tc);
        __syncthreads();

        if (i == j) {
            potrf128_smem(T, s_inv, &s_fail);
            if (s_fail < TB) {
                if (tid == 0) {
                    atomicCAS(a.info, 0, a.row0 + j * TB + s_fail + 1);
                    atomicExch(a.abort, 1);
                }
                return;
            }
            store_lower_tile(T, Gij, a.ld);
            for (int pr = 0; pr < a.n_peers; ++pr) store_lower_tile(T, a.peerA[pr] + (Gij - a.A), a.ld);
            __syncthreads();
            trinv128_smem(T, s_inv, smem + R0_DBL);
            store_lower_tile(T, a.Dinv + (size_t)j * TB * TB, TB);
            for (int pr = 0; pr < a.n_peers; ++pr) store_lower_tile(T, a.peerDinv[pr] + (size_t)j * TB * TB, TB);
        } else {
            if (tid == 0 && !spin_wait(a.ready + (size_t)j * a.ready_ld + j, a.abort)) s_abort = 1;
            __syncthreads();
            if (s_abort) return;
            acc_zero(acc);
            // L_ij[r][c] = sum_k T[r][k] Dinv_j[c][k]
            tile_mainloop<RES_M, STREAM_M>(acc, T, 0, a.Dinv + (size_t)j * TB * TB, TB, 8, smem, &s_abort, NoWait());
            store_tile<false, 1>(acc, Gij, a.ld, tc);
            for (int pr = 0; pr < a.n_peers; ++pr) store_tile<false, 1>(acc, a.peerA[pr] + (Gij - a.A), a.ld, tc);
        }
        if (a.trace && tid == 0) a.trace[4 * (size_t)task_id + 2] = globaltimer_ns();
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            st_release(a.ready + (size_t)i * a.ready_ld + j, 1);
            if (a.trace) a.trace[4 * (size_t)task_id + 3] = globaltimer_ns();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// L^-1 tile-task kernel: X = L^-1 (lower triangular), tasks ordered by anti-diagonal d = i - j.
//   d == 0 : X_jj = Dinv_j (copy)
//   d  > 0 : W = sum_{k=j}^{i-1} L_ik X_kj   (waits on readyX[k][j]);   X_ij = -Dinv_i W
// ---------------------------------------------------------------------------------------------
struct LinvArgs {
    const double* L;
    double* X;
    size_t ld;
    int nb;
    const double* Dinv;
    int* ready;
    int* counter;
    int task_end;
    int* abort;
};

__global__ void __launch_bounds__(NTHREADS, 1) linv_tiles_kernel(LinvArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_task, s_abort, s_upto;
    const int tid = threadIdx.x;
    const TileCoord tc;
    for (;;) {
        if (tid == 0) {
            s_task = atomicAdd(a.counter, 1);
            s_abort = ld_volatile(a.abort) != 0;
            s_upto = 1 << 30;
        }
        __syncthreads();
        int task = s_task;
        if (task >= a.task_end || s_abort) return;
        int d = 0;
        while (task >= a.nb - d) { task -= a.nb - d; ++d; }
        const int j = task, i = j + d;
        double* Xij = a.X + (size_t)j * TB * a.ld + (size_t)i * TB;
        if (d == 0) {
            const double* Dj = a.Dinv + (size_t)j * TB * TB;
            for (int idx = tid; idx < TB * TB / 2; idx += NTHREADS) {
                const int r2 = idx & 63, c = idx >> 6;
                double2 v = __ldcg(reinterpret_cast<const double2*>(Dj + (size_t)c * TB) + r2);
                reinterpret_cast<double2*>(Xij + (size_t)c * a.ld)[r2] = v;
            }
        } else {
            Acc acc;
            acc_zero(acc);
            // i operand: L[i-tile rows, k] streamed M-major starting at tile column j
            const double* Li = a.L + (size_t)j * TB * a.ld + (size_t)i * TB;
            // j operand: X[k, j-tile cols] K-major, element (c,k) at X[(j*128+c)*ld + k], k from j*128
            const double* Xj = a.X + (size_t)j * TB * a.ld + (size_t)j * TB;
            const int upto = ready_prefix(d, [&](int t) { return ld_acquire(a.ready + (size_t)(j + t) * a.nb + j) != 0; }, &s_upto);
            if (!tile_mainloop<STREAM_M, STREAM_K>(acc, Li, a.ld, Xj, a.ld, 8 * upto, smem, &s_abort, NoWait())) return;
            for (int kb = upto; kb < d; ++kb) {
                if (tid == 0 && !spin_wait(a.ready + (size_t)(j + kb) * a.nb + j, a.abort)) s_abort = 1;
                __syncthreads();
                if (s_abort) return;
                if (!tile_mainloop<STREAM_M, STREAM_K>(acc, Li + (size_t)kb * TB * a.ld, a.ld, Xj + (size_t)kb * TB, a.ld, 8, smem,
                                                       &s_abort, NoWait())) return;
            }
            double* W = smem;   // region 0, column-major pitch PM: W[k][c] at c*PM + k
            store_tile<true, 1>(acc, W, PM, tc);
            __syncthreads();
            acc_zero(acc);
            // X_ij[r][c] = -sum_k Dinv_i[r][k] W[k][c]
            tile_mainloop<STREAM_M, RES_K>(acc, a.Dinv + (size_t)i * TB * TB, TB, W, 0, 8, smem, &s_abort, NoWait());
            store_tile<true, -1>(acc, Xij, a.ld, tc);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            st_release(a.ready + (size_t)i * a.nb + j, 1);
        }
    }
}

// Dinv_t = L_tt^-1 for every diagonal tile of an already factorised matrix (model import): one CTA per tile.
__global__ void __launch_bounds__(NTHREADS, 1) dinv_from_l_kernel(const double* L, size_t ld, double* Dinv) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double s_inv[TB];
    const int t = blockIdx.x;
    const double* src = L + (size_t)t * TB * ld + (size_t)t * TB;
    for (int idx = threadIdx.x; idx < TB * TB; idx += NTHREADS) {
        const int r = idx & (TB - 1), c = idx >> 7;
        smem[c * PM + r] = r >= c ? src[(size_t)c * ld + r] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x < TB) s_inv[threadIdx.x] = 1.0 / smem[threadIdx.x * PM + threadIdx.x];
    __syncthreads();
    trinv128_smem(smem, s_inv, smem + R0_DBL);
    __syncthreads();
    store_lower_tile(smem, Dinv + (size_t)t * TB * TB, TB);
}

// ---------------------------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------------------------
static PerDeviceOnce g_tile_attr_done;
static cudaError_t ensure_attrs() {
    const int dev = PerDeviceOnce::current();
    if (g_tile_attr_done.done(dev)) return cudaSuccess;
    {
        cudaError_t e0 = cudaFuncSetAttribute(dinv_from_l_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BYTES);
        if (e0 != cudaSuccess) return e0;
    }
    cudaError_t e = cudaFuncSetAttribute(chol_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(linv_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    g_tile_attr_done.set(dev);
    return cudaSuccess;
}

// sync: the scratch ints are [0]=counter [1]=info [2]=abort followed by nb*nb ready flags.
cudaError_t launch_cholesky(double* A, size_t ld, int nb, double* Dinv, int* scratch, int num_sms, int serial,
                            cudaStream_t st, long long* trace, const CholPeers* peers) {
    cudaError_t e = ensure_attrs();
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(scratch, 0, sizeof(int) * (4 + (size_t)nb * nb), st);
    if (e != cudaSuccess) return e;
    CholArgs a;
    a.A = A; a.ld = ld; a.nb = nb; a.Dinv = Dinv; a.trace = trace;
    a.ready_ld = nb; a.ncols = nb; a.row0 = 0;
    a.n_peers = 0;
    for (int pr = 0; pr < MAX_CHOL_PEERS; ++pr) { a.peerA[pr] = nullptr; a.peerDinv[pr] = nullptr; }
    if (peers && !serial) {
        a.n_peers = peers->n < MAX_CHOL_PEERS ? peers->n : MAX_CHOL_PEERS;
        for (int pr = 0; pr < a.n_peers; ++pr) { a.peerA[pr] = peers->L[pr]; a.peerDinv[pr] = peers->Dinv[pr]; }
    }
    a.counter = scratch; a.info = scratch + 1; a.abort = scratch + 2; a.ready = scratch + 4;
    const int ntasks = nb * (nb + 1) / 2;
    a.task_begin = 0;
    if (!serial) {
        a.task_end = ntasks;
        int grid = ntasks < num_sms ? ntasks : num_sms;
        chol_tiles_kernel<<<grid, NTHREADS, TILE_SMEM_BYTES, st>>>(a);
        return cudaGetLastError();
    }
    // Debug mode: tasks are launched in dependency levels (never more than one level per launch), so
    // that no task ever waits on a flag inside a launch: (0,0); then per column j: (j+1,j) | (j+1,j+1) |
    // the rest of column j.
    auto run = [&](int t0, int t1) {
        if (t1 <= t0) return;
        a.task_begin = t0; a.task_end = t1;
        cudaMemsetAsync(a.counter, 0, sizeof(int), st);
        const int cnt = t1 - t0;
        chol_tiles_kernel<<<cnt < num_sms ? cnt : num_sms, NTHREADS, TILE_SMEM_BYTES, st>>>(a);
    };
    run(0, 1);
    int t0 = 1;
    for (int j = 0; j + 1 < nb; ++j) {
        const int len = nb - j;          // tasks contributed by column j
        run(t0, t0 + 1);
        run(t0 + 1, t0 + 2);
        run(t0 + 2, t0 + len);
        t0 += len;
    }
    return cudaGetLastError();
}

size_t ozaki_fit_workspace_bytes(int S, size_t N) { return (size_t)S * N * N + N * sizeof(double) + (N / 128) * (N / 64); }

// INT8-assisted factorisation.  Left-looking by panels of P tile columns:
//   for each panel [c0, c0 + P):   A[c0.., panel] -= L[c0.., 0:c0] L[panel, 0:c0]^T      INT8 tensor cores, exact integer products
//                                  of base-254 digit slices of L, recombined in FP64 (gpr_ozaki.cu MODE 1)
//                                  panel factorised by chol_tiles_kernel (FP64 tensor pipe; k restricted to the panel)
//                                  finished panel cut into int8 digit slices; row i is scaled by the power of two above
//                                  sqrt(K_ii), read from the diagonal beforehand (|L_ik| <= sqrt(K_ii) for an SPD matrix: nothing
//                                  has to be known about L before it exists, and rows of very different size keep all their digits)
// With S = 7 digits of base 254 the dropped part of every product is below 254^-7 = 1.5e-17 of sqrt(K_ii K_jj) per term — the size
// of the FP64 rounding of the same sum (1.1e-16 of the partial sums) — so the factor is as accurate as the all-FP64 one,
// at ~2x the DMMA rate for the (1 - ~1.5 P / nb) share of the flops that lies left of the panels.  Still one producer per
// tile and integer (order-independent) sums: bit-reproducible.  A non-positive pivot is reported like launch_cholesky does
// (the later launches of the sequence then leave at once: the abort flag stays up).
cudaError_t launch_cholesky_int8(double* A, size_t ld, int nb, double* Dinv, int* scratch, int num_sms, cudaStream_t st,
                                 const CholPeers* peers, signed char* Ls, int S, int panel_tiles, int last_tiles, int* ctrl) {
    cudaError_t e = ensure_attrs();
    if (e != cudaSuccess) return e;
    if (panel_tiles < 1) return cudaErrorInvalidValue;
    e = cudaMemsetAsync(scratch, 0, sizeof(int) * (4 + (size_t)nb * nb), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(ctrl, 0, 2 * sizeof(int), st);       // [1]: raised by an INT8 update whose barrier wait timed out (sticky)
    if (e != cudaSuccess) return e;
    const size_t N = (size_t)nb * TB;
    double* row_scale = reinterpret_cast<double*>(Ls + (size_t)S * N * N);          // tail of the workspace (8-byte aligned: N % 128 == 0)
    e = launch_ozaki_diag_scale(A, ld, (int)N, row_scale, st);
    if (e != cudaSuccess) return e;
    // zero-slice map of the slices (one byte per (128-row tile, 64-k block), bit t = slice t holds a nonzero digit): an entry of L
    // is small against its row's scale once a few columns have been eliminated, so its leading digits vanish in whole blocks
    // and the update skips their MMAs
    unsigned char* nz = reinterpret_cast<unsigned char*>(row_scale + N);
    const size_t nz_pitch = N / 64;
    // The last panels have so few row tiles below them that the tile kernel is bound by its dependency chain (~100 us per
    // column) whatever their width: the final `last_tiles` columns are one panel (two INT8 launches and slicing passes fewer;
    // the fit passes 48, GPR_FIT_LAST overrides).
    for (int c0 = 0, ncols = 0; c0 < nb; c0 += ncols) {
        ncols = nb - c0 < panel_tiles ? nb - c0 : panel_tiles;
        if (nb - c0 <= last_tiles) ncols = nb - c0;
        const size_t r0 = (size_t)c0 * TB;
        e = launch_ozaki_syrk_update(Ls, ld, ld * N, S, r0, N, (size_t)ncols * TB, A, ld, row_scale, ctrl, st, nz, nz_pitch);
        if (e != cudaSuccess) return e;
        CholArgs a;
        a.A = A + r0 * ld + r0; a.ld = ld; a.nb = nb - c0; a.Dinv = Dinv + (size_t)c0 * TB * TB; a.trace = nullptr;
        a.ready = scratch + 4 + (size_t)c0 * nb + c0; a.ready_ld = nb; a.ncols = ncols; a.row0 = (int)r0;
        a.counter = scratch; a.info = scratch + 1; a.abort = scratch + 2;
        a.n_peers = 0;
        for (int pr = 0; pr < MAX_CHOL_PEERS; ++pr) { a.peerA[pr] = nullptr; a.peerDinv[pr] = nullptr; }
        if (peers) {
            a.n_peers = peers->n < MAX_CHOL_PEERS ? peers->n : MAX_CHOL_PEERS;
            for (int pr = 0; pr < a.n_peers; ++pr) { a.peerA[pr] = peers->L[pr] + r0 * ld + r0; a.peerDinv[pr] = peers->Dinv[pr] + (size_t)c0 * TB * TB; }
        }
        // tasks of the first ncols columns in chol_task_decode order (the diagonal task of column ncols, which that order places
        // among them, is skipped by the kernel)
        int ntasks = 1;
        for (int j = 0; j < ncols && j < a.nb - 1; ++j) ntasks += a.nb - j;
        a.task_begin = 0; a.task_end = ntasks;
        if (c0 > 0) { e = cudaMemsetAsync(scratch, 0, sizeof(int), st); if (e != cudaSuccess) return e; }
        chol_tiles_kernel<<<ntasks < num_sms ? ntasks : num_sms, NTHREADS, TILE_SMEM_BYTES, st>>>(a);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (c0 + ncols < nb) {
            e = launch_ozaki_slice_lpanel(A, ld, r0, N, (size_t)ncols * TB, row_scale, S, Ls, ld, ld * N, st);
            if (e != cudaSuccess) return e;
            e = launch_ozaki_mask(Ls + r0 * ld + r0, ld, ld * N, S, (int)((N - r0) / TB), ncols * TB / 64, nz + (r0 / TB) * nz_pitch + r0 / 64,
                                  nz_pitch, st);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

cudaError_t launch_dinv_from_l(const double* L, size_t ld, int nb, double* Dinv, cudaStream_t st) {
    cudaError_t e = ensure_attrs();
    if (e != cudaSuccess) return e;
    if (nb <= 0) return cudaSuccess;
    dinv_from_l_kernel<<<nb, NTHREADS, TILE_SMEM_BYTES, st>>>(L, ld, Dinv);
    return cudaGetLastError();
}

cudaError_t launch_linv(const double* L, double* X, size_t ld, int nb, const double* Dinv, int* scratch, int num_sms,
                        cudaStream_t st) {
    cudaError_t e = ensure_attrs();
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(scratch, 0, sizeof(int) * (4 + (size_t)nb * nb), st);
    if (e != cudaSuccess) return e;
    LinvArgs a;
    a.L = L; a.X = X; a.ld = ld; a.nb = nb; a.Dinv = Dinv;
    a.counter = scratch; a.abort = scratch + 2; a.ready = scratch + 4;
    const int ntasks = nb * (nb + 1) / 2;
    a.task_end = ntasks;
    int grid = ntasks < num_sms ? ntasks : num_sms;
    linv_tiles_kernel<<<grid, NTHREADS, TILE_SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace gpr

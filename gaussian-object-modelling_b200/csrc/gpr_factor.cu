// gpr_factor.cu — K2: blocked FP64 Cholesky and the triangular inverse L^-1, each as ONE persistent
// tile-task kernel on the FP64 tensor pipe (DMMA).
//
// Replaces Eigen's unblocked, pivoted, single-threaded LDLT::compute in the reference
// (/root/reference/include/gp_regression/gp_regressor.hpp:161-162, :457-458).  On the SPD inputs of
// every BASELINE config LDLT and LLT agree to cond(K)*eps; a non-positive pivot is
// reported (never NaN), see SURVEY F2.
//
// Algorithm (left-looking by 128x128 tiles, dependency flags in global memory):
//   task (i,j), i >= j, claimed from an atomic counter in column-major order:
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
    int* ready;       // nb*nb flags, [i*nb + j]
    int* counter;     // tasks claimed so far by this launch
    int task_begin;   // tasks [task_begin, task_end) are executed by this launch
    int task_end;
    int* info;        // 0, or 1 + global index of the first non-positive pivot
    int* abort;       // raised on failure so that waiting CTAs leave
    long long* trace; // optional: 4 globaltimer stamps per task (claim, accumulated, solved, published)
};

__global__ void __launch_bounds__(NTHREADS, 1) chol_tiles_kernel(CholArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_task, s_abort, s_fail, s_upto;
    __shared__ double s_inv[TB];
    const int tid = threadIdx.x;
    const TileCoord tc;
    for (;;) {
        if (tid == 0) {
            s_task = a.task_begin + atomicAdd(a.counter, 1);
            s_abort = ld_volatile(a.abort) != 0;
            s_fail = 1 << 20;
            s_upto = 1 << 30;
        }
        __syncthreads();
        int task = s_task;
        if (task >= a.task_end || s_abort) return;
        const int task_id = task;
        int j = 0;
        while (task >= a.nb - j) { task -= a.nb - j; ++j; }
        const int i = j + task;
        if (a.trace && tid == 0) a.trace[4 * (size_t)task_id + 0] = globaltimer_ns();

        Acc acc;
        acc_zero(acc);
        const double* Li = a.A + (size_t)i * TB;     // row panel i, k = 0
        const double* Lj = a.A + (size_t)j * TB;
        // One bulk look at the readiness flags of both row panels: the k-blocks [0, upto) are complete
        // (almost always all but the last one or two), so the mainloop only polls beyond that prefix
        // instead of paying an L2 round trip in front of a barrier at every 128-wide block.
        const int* fi = a.ready + (size_t)i * a.nb;
        const int* fj = a.ready + (size_t)j * a.nb;
        const int upto = ready_prefix(j, [&](int t) { return ld_acquire(fi + t) != 0 && ld_acquire(fj + t) != 0; }, &s_upto);
        auto waitf = [&](int kb) -> bool {
            if (tid != 0 || kb < upto) return true;
            if (!spin_wait(fi + kb, a.abort)) return false;
            if (i != j && !spin_wait(fj + kb, a.abort)) return false;
            return true;
        };
        if (!tile_mainloop<STREAM_M, STREAM_M>(acc, Li, a.ld, Lj, a.ld, 8 * j, smem, &s_abort, waitf)) return;

        if (a.trace && tid == 0) a.trace[4 * (size_t)task_id + 1] = globaltimer_ns();
        double* T = smem;   // region 0, column-major pitch PM
        double* Gij = a.A + (size_t)j * TB * a.ld + (size_t)i * TB;
        residual_to_smem<false>(acc, Gij, a.ld, T, tc);
        __syncthreads();

        if (i == j) {
            potrf128_smem(T, s_inv, &s_fail);
            if (s_fail < TB) {
                if (tid == 0) {
                    atomicCAS(a.info, 0, j * TB + s_fail + 1);
                    atomicExch(a.abort, 1);
                }
                return;
            }
            store_lower_tile(T, Gij, a.ld);
            __syncthreads();
            trinv128_smem(T, s_inv, smem + R0_DBL);
            store_lower_tile(T, a.Dinv + (size_t)j * TB * TB, TB);
        } else {
            if (tid == 0 && !spin_wait(a.ready + (size_t)j * a.nb + j, a.abort)) s_abort = 1;
            __syncthreads();
            if (s_abort) return;
            acc_zero(acc);
            // L_ij[r][c] = sum_k T[r][k] Dinv_j[c][k]
            tile_mainloop<RES_M, STREAM_M>(acc, T, 0, a.Dinv + (size_t)j * TB * TB, TB, 8, smem, &s_abort, NoWait());
            store_tile<false, 1>(acc, Gij, a.ld, tc);
        }
        if (a.trace && tid == 0) a.trace[4 * (size_t)task_id + 2] = globaltimer_ns();
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            st_release(a.ready + (size_t)i * a.nb + j, 1);
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
            auto waitf = [&](int kb) -> bool {
                if (tid != 0 || kb < upto) return true;
                return spin_wait(a.ready + (size_t)(j + kb) * a.nb + j, a.abort);
            };
            if (!tile_mainloop<STREAM_M, STREAM_K>(acc, Li, a.ld, Xj, a.ld, 8 * d, smem, &s_abort, waitf)) return;
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

// ---------------------------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------------------------
static int g_tile_attr_done = 0;
static cudaError_t ensure_attrs() {
    if (g_tile_attr_done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(chol_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(linv_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    g_tile_attr_done = 1;
    return cudaSuccess;
}

// sync: the scratch ints are [0]=counter [1]=info [2]=abort followed by nb*nb ready flags.
cudaError_t launch_cholesky(double* A, size_t ld, int nb, double* Dinv, int* scratch, int num_sms, int serial,
                            cudaStream_t st, long long* trace) {
    cudaError_t e = ensure_attrs();
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(scratch, 0, sizeof(int) * (4 + (size_t)nb * nb), st);
    if (e != cudaSuccess) return e;
    CholArgs a;
    a.A = A; a.ld = ld; a.nb = nb; a.Dinv = Dinv; a.trace = trace;
    a.counter = scratch; a.info = scratch + 1; a.abort = scratch + 2; a.ready = scratch + 4;
    const int ntasks = nb * (nb + 1) / 2;
    a.task_begin = 0;
    if (!serial) {
        a.task_end = ntasks;
        int grid = ntasks < num_sms ? ntasks : num_sms;
        chol_tiles_kernel<<<grid, NTHREADS, TILE_SMEM_BYTES, st>>>(a);
        return cudaGetLastError();
    }
    // Debug mode: one launch for each diagonal task and one for the rest of its column, so that no
    // task ever waits on a flag inside a launch.
    int t0 = 0;
    for (int j = 0; j < nb; ++j) {
        a.task_begin = t0; a.task_end = t0 + 1;
        cudaMemsetAsync(a.counter, 0, sizeof(int), st);
        chol_tiles_kernel<<<1, NTHREADS, TILE_SMEM_BYTES, st>>>(a);
        int rest = nb - j - 1;
        if (rest > 0) {
            a.task_begin = t0 + 1; a.task_end = t0 + 1 + rest;
            cudaMemsetAsync(a.counter, 0, sizeof(int), st);
            chol_tiles_kernel<<<rest < num_sms ? rest : num_sms, NTHREADS, TILE_SMEM_BYTES, st>>>(a);
        }
        t0 += nb - j;
    }
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

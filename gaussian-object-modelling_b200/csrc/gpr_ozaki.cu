// gpr_ozaki.cu — K3'': the variance product V = X K*^T (X = L^-1) on the INT8 tensor cores of sm_100a
// (tcgen05.mma kind::i8, TMEM int32 accumulators, TMA operand feeds), FP64-equivalent by Ozaki-style slicing.
//
// Why: the FP64 tensor pipe (DMMA) is the roof of var_tiles_kernel / var_trsm_kernel (37 TF/s); tcgen05 has no f64
// kind, but it multiplies 8-bit integers EXACTLY into 32-bit accumulators at ~60x that rate.  So, with digits d in base B,
//   X   = 2^(e_i) * sum_t A_t / (F B^t)      (row i scaled by a power of two; A_t int8 digit slices)
//   K*  = 2^(g)   * sum_u B_u / (F B^u)      (one power-of-two scale per batch; B_u int8 digit slices)
//   V_iq = 2^(e_i+g) * sum_l w_l * [ sum_{t+u=l} sum_k A_t[i,k] B_u[q,k] ],   w_l = 1 / (F^2 B^l),   l = 0 .. S-1
// where every bracket is an exact integer computed by int8 MMAs and the S levels are recombined in FP64 in the epilogue,
// which also reduces the squared column norms per 128-row tile exactly like the DMMA kernels
// (partial[row tile][query] -> var_finalize_kernel).  Two digit systems:
//   base 254: F = 127, |digit| <= 127 (7.99 bits per slice); an int32 accumulator holds S * 127^2 * k < 2^31, i.e. k <= 22016 for S = 6;
//   base 128: F = 64,  |digit| <= 64  (7 bits per slice);    S * 64^2 * k < 2^31, i.e. k <= 74752 for S = 7.
// Longer rows (config 5: n = 65536) run the CHUNKED variant of the kernel: the k range of a task is cut into chunks of at
// most that length, after each of which the epilogue drains the int32 accumulators into FP64 running sums (registers; 64-query
// tiles) — still exact integer arithmetic per chunk, one FP64 rounding per chunk and level on top (<= 1e-16 relative).
// Slices beyond level S - 1 are dropped: truncation ~ B^(-S) relative to (row max of X) x (max of K*) x sqrt(k); the slice
// count is chosen against the variance tolerance (profiles/ozaki_slicing_study_r2.json, tools/ozaki_study.py), and every
// call is spot-checked against the FP64 tensor pipe by the caller (gpr_c_api.cu).
//
// Kernel structure (one CTA per SM, persistent over (row tile, query tile) tasks dealt round-robin):
//   * operands: int8 slice tensors [slice][row][k] (k contiguous = K-major), fetched by TMA (3-D boxes
//     {64 k, 128 or BN rows, S slices}, SWIZZLE_64B) into a ring of shared-memory stages, completion on mbarriers;
//   * warp 0: TMA producer; warp 1: one thread issues the tcgen05.mma instructions (M = 128, N = BN = 64 or 80, K = 32 per
//     instruction; all slice pairs of a k-block reuse the stage); tcgen05.commit releases a stage / signals the epilogue;
//   * accumulators: S levels x BN TMEM columns (<= 512), one 128 x BN int32 tile per level;
//   * warps 2..5: epilogue — each thread reads its TMEM lane (row) with tcgen05.ld, recombines in FP64, squares, reduces.
// X is lower triangular: row tile rt only visits k < 128 (rt + 1).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "gpr_common.cuh"
#include "gpr_kernels.h"

namespace gpr {

constexpr int OZ_BM = 128;        // rows of X per task (UMMA M)
constexpr int OZ_BN_MAX = 80;     // queries per task (UMMA N): 64 or 80 (S * BN <= 512 TMEM columns)
constexpr int OZ_BK = 64;         // k per pipeline stage (bytes per row = SWIZZLE_64B span); 2 MMAs of K = 32
constexpr int OZ_TMEM_COLS = 512;
constexpr int OZ_A_SLICE_BYTES = OZ_BM * OZ_BK;     // 8 KB
constexpr long long OZ_TIMEOUT = 4000000000LL;      // cycles (~2 s): a bug must not hang the GPU

struct OzArgs {
    int S, levels;                // slices per operand; levels = pairs (t, u) with t + u < levels are used
    int stages;                   // pipeline depth
    int nrt, nqt;                 // row tiles (128), query tiles (64)
    int tri;                      // 1: row tile rt needs k < 128 (rt + 1) only
    int kblocks;                  // K / 64 (tri == 0)
    const double* row_scale;      // per row of X: 2^(e_i)
    double col_scale;             // 2^(g)
    double* partial;              // [nrt][q_pad]
    size_t q_pad;
    int* ctrl;                    // [0] task counter, [1] abort flag
    int* dbg;                     // optional raw accumulators [S][nrt*128][dbg_ld]
    size_t dbg_ld;
    int gr, gq;                   // co-scheduled group: gr row tiles x gq query tiles
    double wl[8];                 // level weights 1 / (F^2 B^l)
    const unsigned char* nzA;     // optional [nrt][nz_pitch]: bit t set iff slice t of (row tile, 64-wide k-block) has a nonzero digit
    size_t nz_pitch;
    int kchunk;                   // k-blocks (of 64) after which the int32 accumulators are drained into FP64 (they could overflow
                                  // beyond); >= the longest row for the unchunked kernel
    double* C;                    // MODE 1 (symmetric update of the factorisation): C[col * ldc + row] -= col_scale * product, rows of
    size_t ldc;                   // the A operand x rows of the B operand, tiles on or below the diagonal only
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Returns false after ~2 s (a lost arrival must end the kernel, not hang the device).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > OZ_TIMEOUT) return false;
    return true;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of one K-major slice tile written by TMA with SWIZZLE_64B: rows of 64 bytes, 8-row
// groups 512 bytes apart (SBO), leading-dimension field 1 (ignored for swizzled K-major), descriptor version 1
// (sm_100), layout type 4 = SWIZZLE_64B.  byte_off: K offset inside the 64-byte row (0 or 32).
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr, uint32_t byte_off) {
    uint64_t d = (uint64_t)(((smem_addr + byte_off) >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// Instruction descriptor, kind::i8: D = S32 (c_format 2), A and B signed 8-bit (format 1), both K-major, N >> 3, M >> 4.
__device__ __forceinline__ uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- the kernel ---------------------------------------------------------------------------------------------
// Warp roles (192 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, warps 2..5 = epilogue (TMEM lane
// quarter = warp % 4).  Tasks (pairs of row tiles x one query tile) are dealt round-robin (task = blockIdx.x + i * gridDim.x),
// so every role derives the same task sequence without communication.  With N = 64 a tcgen05.mma lasts only ~32-48 cycles: the
// issuing thread's instruction count per MMA is what limits the rate, hence the fully unrolled, descriptor-incrementing
// issue loop (template on the slice count) and the separate producer warp.
constexpr int OZ_NTHREADS = 192;

// MODE 0: variance (squared column norms of the product per row tile -> partial).  MODE 1: C -= product (the trailing update
// of the INT8-assisted Cholesky, launch_ozaki_syrk_update): both operands are row ranges of the same slice tensor, tasks above
// the diagonal are skipped, the epilogue carries -C / scale in its FP64 running sums (loaded before the first wait).
template <int S, int BN, bool CHUNKED, int MODE>
__global__ void __launch_bounds__(OZ_NTHREADS, 1)
ozaki_var_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OzArgs a) {
    extern __shared__ __align__(1024) uint8_t oz_smem[];
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], accum_full, accum_empty;
    __shared__ uint32_t s_tmem;
    __shared__ double sred[4][OZ_BN_MAX];
    static_assert(S * BN <= OZ_TMEM_COLS && BN % 16 == 0 && BN <= OZ_BN_MAX, "TMEM budget");
    constexpr int OZ_BN = BN;
    constexpr int OZ_B_SLICE_BYTES = BN * OZ_BK;

    constexpr int LEVELS = S;
    constexpr uint32_t STAGE_BYTES = (uint32_t)S * (OZ_A_SLICE_BYTES + OZ_B_SLICE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)oz_smem + 1023) & ~(uintptr_t)1023);
    const int stages = a.stages;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&accum_full, 1);
        mbar_init(&accum_empty, 4);                       // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(OZ_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    const int per_group = a.gr * a.gq;
    const int qgroups = (a.nqt + a.gq - 1) / a.gq;
    // A task is a PAIR of row tiles (nrt-1-p, p) for one query tile, done one after the other (h = 0, 1): with a lower
    // triangular A the two k ranges add up to nrt + 1 blocks of 128 for every pair, so all tasks have the same length and
    // the CTAs that run together stay in step — which is what lets them share operand tiles through L2.
    const int npairs = (a.nrt + 1) / 2;
    const int rgroups = (npairs + a.gr - 1) / a.gr;
    // MODE 1: every task has the same length (k < r0), so the valid (row tile, query tile) pairs are simply enumerated compactly
    // and dealt round-robin: first the row tiles below the diagonal block (nqt tasks each, bottom up), then the diagonal block's
    // row tiles rt < nqt * BN / 128 with their 2 (rt + 1) tiles on or below the diagonal.
    const int dblk = MODE == 1 ? a.nqt * BN / OZ_BM : 0;                            // row tiles of the diagonal block
    const int full_tasks = MODE == 1 ? (a.nrt - dblk) * a.nqt : 0;
    const int ntasks = MODE == 1 ? full_tasks + dblk * (dblk + 1) * (OZ_BM / BN) / 2 : rgroups * qgroups * per_group;
    auto decode = [&](int task, int h, int& rt, int& qt) {
        if (MODE == 1) {
            if (h) return false;
            if (task < full_tasks) { rt = a.nrt - 1 - task / a.nqt; qt = task % a.nqt; return true; }
            int t = task - full_tasks;
            rt = dblk - 1;
            while (t >= (OZ_BM / BN) * (rt + 1)) { t -= (OZ_BM / BN) * (rt + 1); --rt; }
            qt = t;
            return true;
        }
        const int g = task / per_group, w = task % per_group;
        const int rg = g / qgroups, qg = g % qgroups;
        const int pr = rg * a.gr + w % a.gr;
        qt = qg * a.gq + w / a.gr;
        rt = h == 0 ? a.nrt - 1 - pr : pr;
        return pr < npairs && qt < a.nqt && !(h == 1 && pr == a.nrt - 1 - pr);       // odd nrt: the middle row tile is its own pair
    };

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        uint32_t j = 0;
        bool ok = true;
        for (int it = 0; ok; ++it) {
            const int task = blockIdx.x + (it >> 1) * gridDim.x;
            if (task >= ntasks) break;
            int rt, qt;
            if (!decode(task, it & 1, rt, qt)) continue;
            const int nkb = a.tri ? 2 * (rt + 1) : a.kblocks;
            for (int kb = 0; kb < nkb; ++kb, ++j) {
                const uint32_t s = j % (uint32_t)stages, u = j / (uint32_t)stages;
                if (u > 0 && !mbar_wait(&empty_bar[s], (u - 1) & 1)) { ok = false; break; }
                uint8_t* sa = base + (size_t)s * STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                tma_load_3d(sa, &tmA, &full_bar[s], kb * OZ_BK, rt * OZ_BM, 0);
                tma_load_3d(sa + S * OZ_A_SLICE_BYTES, &tmB, &full_bar[s], kb * OZ_BK, qt * OZ_BN, 0);
            }
        }
        if (!ok) atomicExch(a.ctrl + 1, 1);
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        constexpr uint32_t DESC_HI = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);      // SBO | version 1 | SWIZZLE_64B
        uint32_t j = 0, tcount = 0;
        bool ok = true;
        for (int it = 0; ok; ++it) {
            const int task = blockIdx.x + (it >> 1) * gridDim.x;
            if (task >= ntasks) break;
            int rt, qt;
            if (!decode(task, it & 1, rt, qt)) continue;
            const int nkb = a.tri ? 2 * (rt + 1) : a.kblocks;
            // Digit slices of A that are entirely zero in a (row tile, k-block) — far from the diagonal the entries of L^-1 are
            // small against their row maximum, so the leading digits vanish (36 % of the blocks at n = 16384) — contribute
            // nothing: their MMAs are skipped.  The first k-block of a task never skips (it zero-initialises every level).
            // Long rows (CHUNKED): the k range is cut into chunks after each of which the epilogue drains the int32 accumulators
            // into FP64, so that no accumulator can overflow; each chunk is one accumulate / commit / drain cycle.
            const unsigned char* nzrow = a.nzA ? a.nzA + (size_t)rt * a.nz_pitch : nullptr;
            const int kch = CHUNKED ? a.kchunk : nkb;
          for (int c0 = 0; c0 < nkb && ok; c0 += kch) {
            const int c1 = c0 + kch < nkb ? c0 + kch : nkb;
            if (tcount > 0 && !mbar_wait(&accum_empty, (tcount - 1) & 1)) { ok = false; break; }   // epilogue has drained TMEM
            tc_fence_after();
            uint32_t mask_next = 0xFFu;
            for (int kb = c0; kb < c1; ++kb, ++j) {
                const uint32_t s = j % (uint32_t)stages, u = j / (uint32_t)stages;
                const uint32_t mask = mask_next;
                mask_next = (nzrow && kb + 1 < c1) ? (uint32_t)__ldg(nzrow + kb + 1) : 0xFFu;
                if (!mbar_wait(&full_bar[s], u & 1)) { ok = false; break; }
                tc_fence_after();
                const uint32_t sa = smem_u32(base + (size_t)s * STAGE_BYTES);
                const uint32_t a_lo = (((sa >> 4) & 0x3FFFu) | (1u << 16));
                const uint32_t b_lo = a_lo + (uint32_t)(S * OZ_A_SLICE_BYTES >> 4);
                const uint32_t keep = kb > c0 ? 1u : 0u;
                // Slice t of A meets slices u = 0 .. S-1-t of B, whose products belong to the levels t .. S-1: the B slices lie
                // back to back in the stage (BN rows of 64 bytes each, a multiple of the swizzle atom) and the level accumulators
                // back to back in TMEM, so ONE instruction with N = cnt * BN covers cnt slices of B at once (cnt * BN <= 256).
                // Against one instruction per (t, u) pair that is the same tensor work with far fewer reads of the A tile from
                // shared memory (S = 6: 9 instead of 21 per 32-k step) — the operand path, not the tensor pipe, was the busiest unit.
                constexpr int MAXC = 256 / OZ_BN;
#pragma unroll
                for (int t = 0; t < S; ++t) {
                    if (!((mask >> t) & 1u)) continue;
#pragma unroll
                    for (int u0 = 0; u0 < S - t; u0 += MAXC) {
                        const int cnt = (S - t - u0) < MAXC ? (S - t - u0) : MAXC;
                        const uint32_t idesc_c = umma_idesc_i8(OZ_BM, cnt * OZ_BN);
#pragma unroll
                        for (int ks = 0; ks < OZ_BK / 32; ++ks) {
                            const uint64_t da = ((uint64_t)DESC_HI << 32) | (a_lo + (uint32_t)(t * (OZ_A_SLICE_BYTES >> 4) + 2 * ks));
                            const uint64_t db = ((uint64_t)DESC_HI << 32) | (b_lo + (uint32_t)(u0 * (OZ_B_SLICE_BYTES >> 4) + 2 * ks));
                            umma_i8(tmem + (uint32_t)((t + u0) * OZ_BN), da, db, idesc_c, (t == 0 && ks == 0) ? keep : 1u);
                        }
                    }
                }
                umma_commit(&empty_bar[s]);                            // stage free once these MMAs have read it
            }
            if (!ok) break;
            umma_commit(&accum_full);                                  // all MMAs of the chunk done -> epilogue
            ++tcount;
          }
        }
        if (!ok) atomicExch(a.ctrl + 1, 1);
    } else if (warp >= 2) {
        // ===== epilogue: thread owns TMEM lane = its row of the tile =====
        const int q4 = warp & 3;                                       // TMEM lane quarter this warp may read
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        uint32_t tcount = 0;
        bool ok = true;
        for (int it = 0; ok; ++it) {
            const int task = blockIdx.x + (it >> 1) * gridDim.x;
            if (task >= ntasks) break;
            int rt, qt;
            if (!decode(task, it & 1, rt, qt)) continue;
            const int row = rt * OZ_BM + q4 * 32 + lane;
            // MODE 1: both operands carry per-row power-of-two scales (row_scale indexed from the panel's first row)
            const double rs = MODE == 1 ? a.row_scale[row] : a.row_scale[row] * a.col_scale;
            if constexpr (!CHUNKED) {
                if (!mbar_wait(&accum_full, tcount & 1)) { ok = false; break; }
                ++tcount;
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < OZ_BN / 16; ++c) {
                    double acc[16];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.0;
#pragma unroll
                    for (int l = LEVELS - 1; l >= 0; --l) {
                        uint32_t v[16];
                        tmem_ld16(tmem + lane_base + (uint32_t)(l * OZ_BN + 16 * c), v);
                        const double wl = a.wl[l];
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) acc[jj] = fma((double)(int)v[jj], wl, acc[jj]);
                        if (a.dbg) {
                            int* o = a.dbg + ((size_t)l * a.nrt * OZ_BM + row) * a.dbg_ld;
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) {
                                const size_t col = (size_t)qt * OZ_BN + 16 * c + jj;
                                if (col < a.dbg_ld) o[col] = (int)v[jj];
                            }
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        const double vv = acc[jj] * rs;
                        double sv = vv * vv;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
                        if (lane == 0) sred[q4][16 * c + jj] = sv;
                    }
                }
                // TMEM is drained: the MMA issuer may start the next task
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty)) : "memory");
            } else {
                // one drain per k-chunk into FP64 running sums (registers), squares and reduction after the last chunk
                const int nkb = a.tri ? 2 * (rt + 1) : a.kblocks;
                double vacc[OZ_BN];
                double* crow = MODE == 1 ? a.C + (size_t)qt * OZ_BN * a.ldc + row : nullptr;
                const double* cs = MODE == 1 ? a.row_scale + (size_t)qt * OZ_BN : nullptr;      // scales of this tile's columns
                if constexpr (MODE == 1) {
                    const double irs = -1.0 / rs;                               // powers of two: exact
#pragma unroll
                    for (int jj = 0; jj < OZ_BN; ++jj) vacc[jj] = __ldcg(crow + (size_t)jj * a.ldc) * (irs / __ldg(cs + jj));
                } else {
#pragma unroll
                    for (int jj = 0; jj < OZ_BN; ++jj) vacc[jj] = 0.0;
                }
                for (int c0 = 0; c0 < nkb && ok; c0 += a.kchunk) {
                    if (!mbar_wait(&accum_full, tcount & 1)) { ok = false; break; }
                    ++tcount;
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < OZ_BN / 16; ++c) {
                        double acc[16];
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.0;
#pragma unroll
                        for (int l = LEVELS - 1; l >= 0; --l) {
                            uint32_t v[16];
                            tmem_ld16(tmem + lane_base + (uint32_t)(l * OZ_BN + 16 * c), v);
                            const double wl = a.wl[l];
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) acc[jj] = fma((double)(int)v[jj], wl, acc[jj]);
                            if (a.dbg) {                                        // debug dump: the accumulators are SUMMED over the chunks
                                int* o = a.dbg + ((size_t)l * a.nrt * OZ_BM + row) * a.dbg_ld;
#pragma unroll
                                for (int jj = 0; jj < 16; ++jj) {
                                    const size_t col = (size_t)qt * OZ_BN + 16 * c + jj;
                                    if (col < a.dbg_ld) o[col] = (c0 == 0 ? 0 : o[col]) + (int)v[jj];
                                }
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) vacc[16 * c + jj] += acc[jj];
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&accum_empty)) : "memory");
                }
                if (!ok) break;
                if constexpr (MODE == 1) {
#pragma unroll
                    for (int jj = 0; jj < OZ_BN; ++jj) crow[(size_t)jj * a.ldc] = -(vacc[jj] * (rs * __ldg(cs + jj)));
                    continue;
                } else {
#pragma unroll
                    for (int jj = 0; jj < OZ_BN; ++jj) {
                        const double vv = vacc[jj] * rs;
                        double sv = vv * vv;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
                        if (lane == 0) sred[q4][jj] = sv;
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");               // the four epilogue warps
            const int et = tid - 64;
            if (et < OZ_BN && (size_t)qt * OZ_BN + et < a.q_pad)
                a.partial[(size_t)rt * a.q_pad + (size_t)qt * OZ_BN + et] = (sred[0][et] + sred[1][et]) + (sred[2][et] + sred[3][et]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (!ok) atomicExch(a.ctrl + 1, 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(OZ_TMEM_COLS) : "memory");
}

// ---- slicing kernels ---------------------------------------------------------------------------------------
// Row maxima of |X| over the lower triangle (X column-major, ld): bits of the (non-negative) doubles order like
// integers, so atomicMax on the bit pattern is exact.
__global__ void __launch_bounds__(256) oz_rowmax_kernel(const double* __restrict__ X, size_t ld, int n, unsigned long long* rowmax) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int k0 = blockIdx.y * 256;
    if (i >= n) return;
    double m = 0.0;
    const int k1 = min(k0 + 256, i + 1);
    for (int k = k0; k < k1; ++k) m = fmax(m, fabs(X[(size_t)k * ld + i]));
    if (m > 0.0) atomicMax(rowmax + i, (unsigned long long)__double_as_longlong(m));
}
// row_scale[i] = 2^ceil(log2(max)) (>= max, a power of two; 1 for an empty row)
__global__ void oz_rowscale_kernel(const unsigned long long* rowmax, int n, double* row_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = __longlong_as_double((long long)rowmax[i]);
    int e = 0;
    if (m > 0.0) { frexp(m, &e); }                       // m = f * 2^e, f in [0.5, 1)  ->  2^e >= m
    row_scale[i] = ldexp(1.0, e);
}
// One value -> S signed digits in base B with first scale F (F = B / 2): v in [-1, 1];  d_0 = rint(F v), r = F v - d_0 in
// [-1/2, 1/2];  d_t = rint(B r) in [-F, F], r = B r - d_t ...   (B = 128, F = 64 or B = 254, F = 127: all digits fit int8)
// Slices of a column-major FP64 matrix M (element (row r, col c) at M[c*ld + r]) into K-major int8 tensors
// out[t][r][c] (c contiguous, row pitch out_ld, slice pitch out_slice): 64 x 64 tiles through shared memory so that both the
// FP64 reads (along r) and the byte writes (along c) are coalesced.  scale_rows: per-row power-of-two scale (X), or
// nullptr with one common scale inv_scale (the K* panel).  tri: entries with c > r are structural zeros.
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ M, size_t ld, int rows, int cols,
                                                        const double* __restrict__ scale_rows, double inv_scale, int tri,
                                                        int S, double first, double base, signed char* __restrict__ out, size_t out_ld,
                                                        size_t out_slice, int fill_upper) {
    // one buffer, two uses: the FP64 tile (transposed on the way in), then — once every thread holds its 16 values in
    // registers — the digit bytes [slice][row][64 k], which leave as 16-byte stores (byte stores to global memory made this
    // kernel latency-bound: 1.5 ms per K* panel at 35 % of the DRAM bandwidth)
    __shared__ __align__(16) double tile[64 * 65];
    signed char* dig = reinterpret_cast<signed char*>(tile);
    const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    if (tri && c0 > r0 + 63 && !fill_upper) return;      // block entirely above the diagonal: left zero by the memset (or, with
                                                         // fill_upper, written as zeros: the buffer is not cleared beforehand)
    for (int e = threadIdx.x; e < 64 * 64; e += 256) {
        const int rr = e & 63, cc = e >> 6;
        const int r = r0 + rr, c = c0 + cc;
        tile[rr * 65 + cc] = (r < rows && c < cols && !(tri && c > r)) ? M[(size_t)c * ld + r] : 0.0;
    }
    __syncthreads();
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int e = threadIdx.x + 256 * i, cc = e & 63, rr = e >> 6;
        const int r = r0 + rr;
        const double sc = scale_rows ? (r < rows ? 1.0 / scale_rows[r] : 0.0) : inv_scale;
        v[i] = tile[rr * 65 + cc] * sc;
    }
    __syncthreads();
    // digit t of all 16 values before digit t + 1 of any: 16 independent dependency chains in flight per thread
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= first;
    // rint() and the conversion to int without the slow FP64 conversion instructions: v + 1.5 * 2^52 rounds v to the nearest
    // integer (ties to even, like rint) and leaves it, as a two's complement number, in the low word of the sum (|v| < 2^31).
    constexpr double MAGIC = 6755399441055744.0;
    for (int t = 0; t < S; ++t) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double m = __dadd_rn(v[i], MAGIC);
            const double it = __dsub_rn(m, MAGIC);
            dig[t * (64 * 64) + threadIdx.x + 256 * i] = (signed char)__double2loint(m);      // index = rr * 64 + cc
            v[i] = (v[i] - it) * base;
        }
    }
    __syncthreads();
    // [slice][rr][64 bytes] -> global rows of 64 bytes, 16 bytes per thread (rows / columns past the matrix are not written)
    const bool full = r0 + 64 <= rows && c0 + 64 <= cols && (out_ld % 16 == 0) && (out_slice % 16 == 0) && ((reinterpret_cast<uintptr_t>(out) + c0) % 16 == 0);
    if (full) {
        for (int e = threadIdx.x; e < S * 64 * 4; e += 256) {
            const int q = e & 3, rr = (e >> 2) & 63, t = e >> 8;
            const uint4 w = *reinterpret_cast<const uint4*>(dig + (t * 64 + rr) * 64 + 16 * q);
            *reinterpret_cast<uint4*>(out + (size_t)t * out_slice + (size_t)(r0 + rr) * out_ld + c0 + 16 * q) = w;
        }
    } else {
        for (int e = threadIdx.x; e < S * 64 * 64; e += 256) {
            const int cc = e & 63, rr = (e >> 6) & 63, t = e >> 12;
            if (r0 + rr < rows && c0 + cc < cols) out[(size_t)t * out_slice + (size_t)(r0 + rr) * out_ld + c0 + cc] = dig[e];
        }
    }
}

// nz[rt * pitch + kb] bit t = slice t of rows [128 rt, 128 rt + 128) x k [64 kb, 64 kb + 64) holds a nonzero digit.
// One CTA per (k-block, row tile); slices [t][row][k] with row pitch a_pitch and slice pitch a_slice.
__global__ void __launch_bounds__(256) oz_mask_kernel(const signed char* __restrict__ As, size_t a_pitch, size_t a_slice, int S,
                                                       unsigned char* __restrict__ nz, size_t pitch) {
    __shared__ unsigned int sm;
    const int kb = blockIdx.x, rt = blockIdx.y;
    if (threadIdx.x == 0) sm = 0u;
    __syncthreads();
    unsigned int bits = 0u;
    for (int t = 0; t < S; ++t) {
        const signed char* base = As + (size_t)t * a_slice + (size_t)rt * OZ_BM * a_pitch + (size_t)kb * OZ_BK;
        bool any = false;
        for (int e = threadIdx.x; e < OZ_BM * (OZ_BK / 16); e += 256) {            // 128 rows x 4 chunks of 16 bytes
            const int r = e >> 2, c = e & 3;
            const uint4 v = *reinterpret_cast<const uint4*>(base + (size_t)r * a_pitch + 16 * c);
            any = any || (v.x | v.y | v.z | v.w) != 0u;
        }
        if (any) bits |= 1u << t;
    }
    if (bits) atomicOr(&sm, bits);
    __syncthreads();
    if (threadIdx.x == 0) nz[(size_t)rt * pitch + kb] = (unsigned char)sm;
}

cudaError_t launch_ozaki_mask(const signed char* As, size_t a_pitch, size_t a_slice, int S, int nrt, int kblocks, unsigned char* nz,
                              size_t pitch, cudaStream_t st) {
    if (nrt <= 0 || kblocks <= 0) return cudaSuccess;
    oz_mask_kernel<<<dim3((unsigned)kblocks, (unsigned)nrt), 256, 0, st>>>(As, a_pitch, a_slice, S, nz, pitch);
    return cudaGetLastError();
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// [slice][row][k] int8, k contiguous with row pitch `pitch` bytes and slice pitch `slice_pitch`; box {64, box_rows, S}.
static cudaError_t make_map(CUtensorMap* tm, const signed char* ptr, size_t k_extent, size_t rows, int S, size_t pitch,
                            size_t slice_pitch, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    cuuint64_t dims[3] = {(cuuint64_t)k_extent, (cuuint64_t)rows, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)slice_pitch};
    cuuint32_t box[3] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, (cuuint32_t)S};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<signed char*>(ptr), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// Pipeline stages that fit the shared-memory budget.
int ozaki_stages(int S, int BN) {
    const int stage = S * (OZ_A_SLICE_BYTES + BN * OZ_BK);
    int st = (int)((220 * 1024) / stage);
    return st > 6 ? 6 : (st < 1 ? 1 : st);
}
int ozaki_tile_n(int S) { return S * 80 <= OZ_TMEM_COLS ? 80 : 64; }
// Largest k extent for which every level accumulator stays below 2^31 in the worst case.
// Can the engine run S slices over k_extent columns?  Within ozaki_max_k always; beyond it only the slice counts the
// k-chunked kernel is instantiated for.
bool ozaki_supported(int S, int base254, long long k_extent) {
    if (S < 1 || S > 8) return false;
    return k_extent <= ozaki_max_k(S, base254) || S <= 2 || S >= 6;
}

long long ozaki_max_k(int S, int base254) {
    const long long d = base254 ? 127 : 64;
    return ((1LL << 31) - 1) / ((long long)S * d * d);
}

template <int S, int BN, bool CHUNKED, int MODE = 0>
static cudaError_t launch_oz(const CUtensorMap& tmA, const CUtensorMap& tmB, const OzArgs& a, int grid, size_t smem, cudaStream_t st) {
    static PerDeviceOnce attr_done;
    const int cur = PerDeviceOnce::current();
    if (!attr_done.done(cur)) {
        cudaError_t e = cudaFuncSetAttribute(ozaki_var_kernel<S, BN, CHUNKED, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done.set(cur);
    }
    ozaki_var_kernel<S, BN, CHUNKED, MODE><<<grid, OZ_NTHREADS, smem, st>>>(tmA, tmB, a);
    return cudaGetLastError();
}

// Launch over slice tensors that are already built.  As: [S][a_rows][k] (k pitch a_pitch, a_rows = 128 nrt); Bs: [S][b_rows][k]
// (k pitch b_pitch).  base254: digit system (selects the level weights).  partial: [nrt][q_pad].  ctrl: 2 ints.
// dbg: optional raw accumulators [S][a_rows][dbg_ld].
cudaError_t launch_ozaki_product(const signed char* As, size_t a_pitch, size_t a_slice, int nrt, const signed char* Bs, size_t b_pitch,
                                 size_t b_slice, size_t b_rows, int q, size_t q_pad, size_t k_extent, int tri, int S, int base254,
                                 const double* row_scale, double col_scale, double* partial, int* ctrl, int* dbg, size_t dbg_ld,
                                 cudaStream_t st, const unsigned char* nzA, size_t nz_pitch) {
    if (S < 1 || S > 8) return cudaErrorInvalidValue;
    // Beyond ozaki_max_k an int32 accumulator could overflow: the chunked kernel drains the accumulators into FP64 every
    // kchunk k-blocks (64-query tiles: the FP64 running sums live in the epilogue threads' registers).
    // GPR_OZ_KCHUNK=<even number of k-blocks> forces the chunked kernel with that chunk length (tests of the chunk logic).
    const char* kc_env = getenv("GPR_OZ_KCHUNK");
    const int kc_forced = kc_env ? (atoi(kc_env) & ~1) : 0;
    const bool chunked = (long long)k_extent > ozaki_max_k(S, base254) || (kc_forced > 0 && (S <= 2 || S >= 6));
    const int BN = chunked ? 64 : ozaki_tile_n(S);
    const size_t smem = (size_t)220 * 1024 + 1024;
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map(&tmA, As, k_extent, (size_t)nrt * OZ_BM, S, a_pitch, a_slice, OZ_BM);
    if (e != cudaSuccess) return e;
    e = make_map(&tmB, Bs, k_extent, b_rows, S, b_pitch, b_slice, BN);
    if (e != cudaSuccess) return e;
    OzArgs a;
    a.S = S; a.levels = S; a.stages = ozaki_stages(S, BN);
    a.nrt = nrt; a.nqt = (int)((q + BN - 1) / BN); a.tri = tri; a.kblocks = (int)(k_extent / OZ_BK);
    a.row_scale = row_scale; a.col_scale = col_scale; a.partial = partial; a.q_pad = q_pad; a.ctrl = ctrl; a.dbg = dbg; a.dbg_ld = dbg_ld;
    static const bool noskip = getenv("GPR_OZ_NOSKIP") && atoi(getenv("GPR_OZ_NOSKIP")) != 0;      // A/B switch for measurements
    a.nzA = noskip ? nullptr : nzA; a.nz_pitch = nz_pitch;
    a.C = nullptr; a.ldc = 0;
    a.kchunk = chunked ? (int)((ozaki_max_k(S, base254) / OZ_BK) & ~1LL) : (1 << 30);
    if (chunked && kc_forced > 0 && kc_forced < a.kchunk) a.kchunk = kc_forced;
    const double F = base254 ? 127.0 : 64.0, B = base254 ? 254.0 : 128.0;
    double w = 1.0 / (F * F);
    for (int l = 0; l < 8; ++l) { a.wl[l] = w; w /= B; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // co-resident tasks form a (gr row-tile pairs) x (gq query tiles) rectangle: an A slice tile is then shared by gq CTAs and a B
    // slice tile by gr CTAs through L2 (GPR_OZ_GR overrides gr; sweep in profiles/ozaki_group_sweep_r2.json)
    static const int gr_env = getenv("GPR_OZ_GR") ? atoi(getenv("GPR_OZ_GR")) : 0;
    a.gr = gr_env >= 1 && gr_env <= 64 ? gr_env : 10;         // 10 pairs x 14 query tiles; with paired row tiles 3..12 are within 6 % (41.4 - 44.0 ms per batch at n = 16384)
    a.gq = sms / a.gr > 0 ? sms / a.gr : 1;
    const int npairs = (a.nrt + 1) / 2;
    if (a.gq > a.nqt) a.gq = a.nqt;
    if (a.gr > npairs) a.gr = npairs;
    e = cudaMemsetAsync(ctrl, 0, 2 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    const int tasks = ((npairs + a.gr - 1) / a.gr) * ((a.nqt + a.gq - 1) / a.gq) * a.gr * a.gq;
    const int grid = tasks < sms ? tasks : sms;
    if (chunked) {
        switch (S) {
            case 1: return launch_oz<1, 64, true>(tmA, tmB, a, grid, smem, st);       // (self-test of the chunk logic)
            case 2: return launch_oz<2, 64, true>(tmA, tmB, a, grid, smem, st);
            case 6: return launch_oz<6, 64, true>(tmA, tmB, a, grid, smem, st);
            case 7: return launch_oz<7, 64, true>(tmA, tmB, a, grid, smem, st);
            case 8: return launch_oz<8, 64, true>(tmA, tmB, a, grid, smem, st);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (S) {
        case 1: return launch_oz<1, 80, false>(tmA, tmB, a, grid, smem, st);
        case 2: return launch_oz<2, 80, false>(tmA, tmB, a, grid, smem, st);
        case 3: return launch_oz<3, 80, false>(tmA, tmB, a, grid, smem, st);
        case 4: return launch_oz<4, 80, false>(tmA, tmB, a, grid, smem, st);
        case 5: return launch_oz<5, 80, false>(tmA, tmB, a, grid, smem, st);
        case 6: return launch_oz<6, 80, false>(tmA, tmB, a, grid, smem, st);
        case 7: return launch_oz<7, 64, false>(tmA, tmB, a, grid, smem, st);
        default: return launch_oz<8, 64, false>(tmA, tmB, a, grid, smem, st);
    }
}

// Slices of X = L^-1 (lower triangular, column-major n x n with leading dimension ld; rows padded to 128 nrt):
// Xs[t][i][k], row pitch = slice row count = ld.  rowmax: ld 8-byte words of scratch.
cudaError_t launch_ozaki_slice_x(const double* X, size_t ld, int n_rows, int S, int base254, signed char* Xs, double* row_scale,
                                 unsigned long long* rowmax, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(rowmax, 0, ld * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(Xs, 0, (size_t)S * ld * ld, st);
    if (e != cudaSuccess) return e;
    oz_rowmax_kernel<<<dim3((n_rows + 255) / 256, (n_rows + 255) / 256), 256, 0, st>>>(X, ld, n_rows, rowmax);
    oz_rowscale_kernel<<<(int)((ld + 255) / 256), 256, 0, st>>>(rowmax, (int)ld, row_scale);
    oz_slice_kernel<<<dim3((n_rows + 63) / 64, (n_rows + 63) / 64), 256, 0, st>>>(X, ld, n_rows, n_rows, row_scale, 0.0, 1, S,
                                                                                base254 ? 127.0 : 64.0, base254 ? 254.0 : 128.0, Xs, ld, ld * ld, 0);
    return cudaGetLastError();
}

// Slices of the K* panel (element (query c, point k) at panel[k*panel_ld + c]) -> Ks[u][query][k] with k pitch k_pitch.
cudaError_t launch_ozaki_slice_panel(const double* panel, size_t panel_ld, int q, int n_k, double inv_scale, int S, int base254,
                                     signed char* Ks, size_t k_pitch, size_t q_pad, cudaStream_t st) {
    // here the "rows" of the slice tensor are the queries and its "columns" the points: M(row = query, col = k) = panel[k*ld + query]
    oz_slice_kernel<<<dim3((q + 63) / 64, (n_k + 63) / 64), 256, 0, st>>>(panel, panel_ld, q, n_k, nullptr, inv_scale, 0, S,
                                                                        base254 ? 127.0 : 64.0, base254 ? 254.0 : 128.0, Ks, k_pitch,
                                                                        q_pad * k_pitch, 0);
    return cudaGetLastError();
}

// ---- INT8-assisted Cholesky (gpr_factor.cu: launch_cholesky_int8) ---------------------------------------------------------
// Per-row scales of the factor-to-be: row_scale[i] = the power of two above sqrt(K_ii) (|L_ik| <= sqrt(K_ii) for an SPD matrix;
// 1 for a non-positive diagonal entry — that factorisation fails anyway).  Read from the diagonal BEFORE the factorisation.
__global__ void oz_diag_scale_kernel(const double* __restrict__ A, size_t ld, int n, double* __restrict__ row_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = A[(size_t)i * ld + i];
    int e = 0;
    if (d > 0.0) frexp(sqrt(d), &e);
    row_scale[i] = d > 0.0 ? ldexp(1.0, e) : 1.0;
}
cudaError_t launch_ozaki_diag_scale(const double* A, size_t ld, int n, double* row_scale, cudaStream_t st) {
    oz_diag_scale_kernel<<<(n + 255) / 256, 256, 0, st>>>(A, ld, n, row_scale);
    return cudaGetLastError();
}

// Slices of a finished panel of L: rows [r0, n_rows) x columns [r0, r0 + width) of the column-major factor A (leading dimension
// ld; r0 on the diagonal, so local entries with column > row are structural zeros and are WRITTEN as zeros) into
// Ls[t][row][k] (k pitch `pitch`, slice pitch `slice`), row i divided by row_scale[i] (launch_ozaki_diag_scale).
cudaError_t launch_ozaki_slice_lpanel(const double* A, size_t ld, size_t r0, size_t n_rows, size_t width, const double* row_scale, int S,
                                      signed char* Ls, size_t pitch, size_t slice, cudaStream_t st) {
    if (n_rows <= r0 || width == 0) return cudaSuccess;
    const int rows = (int)(n_rows - r0), cols = (int)width;
    oz_slice_kernel<<<dim3((rows + 63) / 64, (cols + 63) / 64), 256, 0, st>>>(A + r0 * ld + r0, ld, rows, cols, row_scale + r0, 0.0, 1, S,
                                                                            127.0, 254.0, Ls + r0 * pitch + r0, pitch, slice, 1);
    return cudaGetLastError();
}

// C[r0.., r0 .. r0 + width) -= sum_{k < r0} L[row, k] L[col, k] on the INT8 tensor cores, from the slices of the finished
// columns k < r0 (base-254 digits of L[i, :] / row_scale[i]): the left-looking update of the next panel.  Only tiles on or
// below the diagonal.
// nz (optional): the zero-slice map of Ls (launch_ozaki_mask: [row tile][64-k block], rows from 0, pitch nz_pitch) — the
// leading digits of most of L vanish (its entries shrink with the column index while the scale is that of the row).
cudaError_t launch_ozaki_syrk_update(const signed char* Ls, size_t pitch, size_t slice, int S, size_t r0, size_t n_rows, size_t width,
                                     double* A, size_t ld, const double* row_scale, int* ctrl, cudaStream_t st,
                                     const unsigned char* nz, size_t nz_pitch) {
    if (r0 == 0 || n_rows <= r0 || width == 0) return cudaSuccess;
    if (S < 6 || S > 8 || r0 % OZ_BM || n_rows % OZ_BM || width % OZ_BM) return cudaErrorInvalidValue;
    constexpr int BN = 64;
    const size_t smem = (size_t)220 * 1024 + 1024;
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map(&tmA, Ls + r0 * pitch, r0, n_rows - r0, S, pitch, slice, OZ_BM);
    if (e != cudaSuccess) return e;
    e = make_map(&tmB, Ls + r0 * pitch, r0, width, S, pitch, slice, BN);
    if (e != cudaSuccess) return e;
    OzArgs a;
    a.S = S; a.levels = S; a.stages = ozaki_stages(S, BN);
    a.nrt = (int)((n_rows - r0) / OZ_BM); a.nqt = (int)(width / BN); a.tri = 0; a.kblocks = (int)(r0 / OZ_BK);
    a.row_scale = row_scale + r0; a.col_scale = 1.0; a.partial = nullptr; a.q_pad = 0; a.ctrl = ctrl; a.dbg = nullptr; a.dbg_ld = 0;
    static const bool noskip = getenv("GPR_OZ_NOSKIP") && atoi(getenv("GPR_OZ_NOSKIP")) != 0;
    a.nzA = (nz && !noskip) ? nz + (r0 / OZ_BM) * nz_pitch : nullptr; a.nz_pitch = nz_pitch;
    a.C = A + r0 * ld + r0; a.ldc = ld;
    a.kchunk = (int)((ozaki_max_k(S, 1) / OZ_BK) & ~1LL);
    double w = 1.0 / (127.0 * 127.0);
    for (int l = 0; l < 8; ++l) { a.wl[l] = w; w /= 254.0; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    a.gr = 10; a.gq = sms / a.gr > 0 ? sms / a.gr : 1;
    const int npairs = (a.nrt + 1) / 2;
    if (a.gq > a.nqt) a.gq = a.nqt;
    if (a.gr > npairs) a.gr = npairs;
    // ctrl is cleared once by launch_cholesky_int8 (the abort flag of the whole sequence is sticky)
    const int dblk = a.nqt * BN / OZ_BM;
    const int tasks = (a.nrt - dblk) * a.nqt + dblk * (dblk + 1) * (OZ_BM / BN) / 2;
    const int grid = tasks < sms ? tasks : sms;
    switch (S) {
        case 6: return launch_oz<6, 64, true, 1>(tmA, tmB, a, grid, smem, st);
        case 7: return launch_oz<7, 64, true, 1>(tmA, tmB, a, grid, smem, st);
        default: return launch_oz<8, 64, true, 1>(tmA, tmB, a, grid, smem, st);
    }
}

}  // namespace gpr

// gpr_mma.cuh — the FP64 tensor-pipe tile engine shared by the Cholesky, L^-1 and variance kernels.
//
// One CTA (256 threads, 8 warps) accumulates a 128x128 FP64 tile
//        acc[i][j] += sum_k  Aop(i,k) * Bop(j,k)
// with mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4 — on sm_100a every FP64 mma shape lowers to it;
// tcgen05.mma has no f64 kind, so there is no TMEM/UMMA path for this precision).
//
// Operand roles.  The "i operand" supplies the rows of the output tile and feeds the MMA *B*
// fragment; the "j operand" supplies the columns and feeds the MMA *A* fragment.  With that
// assignment the two accumulator registers of a thread are two rows of the same output column,
// and with the row permutation below four consecutive rows: the tile is stored with 32-byte
// contiguous pieces per thread and full 128-byte lines per quarter-warp.
//
// Shared-memory layouts of one k16 stage (all conflict-free for the fragment loads used):
//   M-major  [k][m], pitch PM=132 doubles : fragment loads are LDS.128 of two adjacent m for one k.
//   K-major  [m][k], pitch PK=20 doubles  : fragment loads are LDS.64 (j operand only).
// A resident operand (a full 128x128 tile already in shared memory) uses pitch PM in either layout.
//
// Warp w owns rows i0 = 64*(w&1) .. +63 and columns j0 = 32*(w>>1) .. +31:
//   8 n-tiles (i) x 4 m-tiles (j) of 8x8, i.e. 32 DMMA per k4 step and 6 LDS.128 per k4 step.
// Fragment <-> tile index maps (g = lane>>2, t = lane&3):
//   i operand, pair p (n-tiles 2p,2p+1):  rows  i0 + 16p + 2g + {0,1}   loaded as one LDS.128
//   accumulators of pair p:               rows  i0 + 16p + 4t + {0,1,2,3} =
//                                         acc[mt][2p][0], acc[mt][2p+1][0], acc[mt][2p][1], acc[mt][2p+1][1]
//   j operand M-major, m-tile mt:         col   j0 + 16(mt>>1) + 2g + (mt&1)
//   j operand K-major, m-tile mt:         col   j0 + 8mt + g
#pragma once
#include "gpr_common.cuh"

namespace gpr {

constexpr int KT = 16;                     // k extent of one pipeline stage
constexpr int PM = 132;                    // pitch of M-major stage / resident tiles (doubles)
constexpr int PK = 20;                     // pitch of K-major stage tiles (doubles)
constexpr int STAGES = 4;
constexpr int STAGE_I = KT * PM;           // doubles per i-operand stage (M-major)
constexpr int STAGE_J = TB * PK;           // doubles per j-operand stage (max of both layouts: 2560 >= 2112)
constexpr int R0_DBL = TB * PM;            // region 0: resident tile / i stages   (16896 doubles)
constexpr int R1_DBL = STAGES * STAGE_J;   // region 1: j stages (or i stages when j is resident)
constexpr int SS_STAGES = 6;               // stage buffers per operand when BOTH operands are streamed
constexpr int SS_DBL = SS_STAGES * (STAGE_I + STAGE_J);                              // 28,032 doubles
constexpr size_t TILE_SMEM_BYTES = (size_t)(SS_DBL > R0_DBL + R1_DBL ? SS_DBL : R0_DBL + R1_DBL) * sizeof(double);   // 224,256 B

static_assert(STAGES * STAGE_I <= R0_DBL, "i stages must fit region 0");
static_assert(STAGES * STAGE_I <= R1_DBL, "i stages must fit region 1");

enum OperandMode { STREAM_M = 0, STREAM_K = 1, RES_M = 2, RES_K = 3 };

typedef double Acc[4][8][2];

__device__ __forceinline__ void acc_zero(Acc& acc) {
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 8; ++n) { acc[m][n][0] = 0.0; acc[m][n][1] = 0.0; }
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Warp / lane coordinates used by every tile routine.
struct TileCoord {
    int lane, warp, g, t, i0, j0;
    __device__ __forceinline__ TileCoord() {
        lane = threadIdx.x & 31; warp = threadIdx.x >> 5; g = lane >> 2; t = lane & 3;
        i0 = (warp & 1) * 64; j0 = (warp >> 1) * 32;
    }
    // tile row of accumulator (pair p, e in 0..3)
    __device__ __forceinline__ int row(int p, int e) const { return i0 + 16 * p + 4 * t + e; }
    template <bool JK>
    __device__ __forceinline__ int col(int mt) const {
        return JK ? (j0 + 8 * mt + g) : (j0 + 16 * (mt >> 1) + 2 * g + (mt & 1));
    }
};

// acc element for (mt, pair p, e): e=0 -> [2p][0], 1 -> [2p+1][0], 2 -> [2p][1], 3 -> [2p+1][1]
#define GPR_ACC(acc, mt, p, e) (acc)[mt][2 * (p) + ((e) & 1)][(e) >> 1]

// k4 sub-steps [KS0, KS1) of one k16 stage.  As: i operand, element (k,m) at As[k*pa + m].  Bs: j operand;
// M-major element (k,m) at Bs[k*pb + m]; K-major element (m,k) at Bs[m*pb + k].
template <bool JK, int KS0, int KS1>
__device__ __forceinline__ void compute_ks(Acc& acc, const double* __restrict__ As, int pa,
                                           const double* __restrict__ Bs, int pb, const TileCoord& tc) {
#pragma unroll
    for (int ks = KS0; ks < KS1; ++ks) {
        const int k = 4 * ks + tc.t;
        double b[8], a[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            double2 v = *reinterpret_cast<const double2*>(As + k * pa + tc.i0 + 16 * p + 2 * tc.g);
            b[2 * p] = v.x; b[2 * p + 1] = v.y;
        }
        if (JK) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = Bs[(tc.j0 + 8 * mt + tc.g) * pb + k];
        } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                double2 v = *reinterpret_cast<const double2*>(Bs + k * pb + tc.j0 + 16 * q + 2 * tc.g);
                a[2 * q] = v.x; a[2 * q + 1] = v.y;
            }
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
}

// Per-thread addressing of the asynchronous stage copies.  One k16 stage of an operand is 1024 chunks of
// 16 bytes, 4 per thread, and the stages are issued strictly in order (kk = 0, 1, 2, ...), so the source
// address is a RUNNING pointer that advances by one stage per issue — no multiplications and no fresh
// address registers inside the DMMA loop (recomputing src + kk*stride for every chunk cost ~30 integer
// instructions per stage and made the address updates wait on the copies still reading those registers):
//   M-major source: element (m,k) at src[m + k*ld] -> stage[k*PM + m]; thread owns column k = tid>>4 and
//                   rows m = 2(tid&15) + 32c: its 4 chunks are 256 bytes apart in global AND shared memory
//                   (compile-time immediates on one address register each).
//   K-major source: element (m,k) at src[k + m*ld] -> stage[m*PK + k]; thread owns m = tid>>3 (+32c),
//                   k = 2(tid&7): 4 running pointers (rows are ld apart).
template <bool KMAJOR>
struct StageCopy {
    const double* p[KMAJOR ? 4 : 1];   // source of this thread's chunk(s) in the NEXT stage to be issued
    size_t stage_stride;               // source elements between consecutive k16 stages
    int dst;                           // offset of the first chunk inside a stage buffer
    __device__ __forceinline__ void init(const double* base, size_t ld) {
        const int tid = threadIdx.x;
        if (!KMAJOR) {
            p[0] = base + (size_t)(tid >> 4) * ld + 2 * (tid & 15);
            stage_stride = (size_t)KT * ld;
            dst = (tid >> 4) * PM + 2 * (tid & 15);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) p[c] = base + (size_t)((tid >> 3) + 32 * c) * ld + 2 * (tid & 7);
            stage_stride = KT;
            dst = (tid >> 3) * PK + 2 * (tid & 7);
        }
    }
    // Copies the next TWO stages into s0 and s1.  All address arithmetic is done before the first copy is
    // issued, into the registers read by the copies of the PREVIOUS pair (issued microseconds ago): a
    // pointer update placed right after the copies that read the pointer stalls on their scoreboard.
    __device__ __forceinline__ void issue_pair(double* s0, double* s1) {
        if (!KMAJOR) {
            const double* g0 = p[0];
            const double* g1 = g0 + stage_stride;
            p[0] = g1 + stage_stride;
            double* d0 = s0 + dst;
            double* d1 = s1 + dst;
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async16(d0 + 32 * c, g0 + 32 * c);
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async16(d1 + 32 * c, g1 + 32 * c);
        } else {
            const double* g[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { g[c] = p[c]; p[c] = g[c] + 2 * KT; }
            double* d0 = s0 + dst;
            double* d1 = s1 + dst;
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async16(d0 + c * 32 * PK, g[c]);
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async16(d1 + c * 32 * PK, g[c] + KT);
        }
    }
};

// Pipelined accumulation over nk16 k16-steps (nk16 even).
//   IMODE in {STREAM_M, RES_M}; JMODE in {STREAM_M, STREAM_K, RES_M, RES_K}.
//   Streamed operands: Ag/Bg point at (tile row 0, k = 0) of the operand; lda/ldb are its leading
//   dimensions.  Resident operands: Ag/Bg are shared-memory tiles with pitch PM.
// The k16 stages are processed in pairs: one barrier per 32 k, the copies of a later pair are issued
// in the middle of the current pair's DMMA stream (so their address arithmetic and the barrier do not
// sit in front of a block of tensor instructions).
//   * both operands streamed (Cholesky accumulation, variance product, L^-1 accumulation): SIX stage buffers carved
//     out of the whole shared-memory block = three pairs in flight; a pair is awaited two iterations
//     (~9 us of DMMA work) after it was issued (cp.async.wait_group 1).  With two pairs in flight 3.3 % of
//     the warp samples sat on the copy scoreboard at the top of the loop (ncu r1e).
//   * otherwise (an operand resident in region 0): four stage buffers in region 1 = two pairs in flight.
// All dependency waits of the callers happen OUTSIDE this loop (s_abort / the wait functor are kept in the
// signature for the callers' convenience and are not used).  Always returns true.  Ends with all async
// copies drained and a __syncthreads().
template <int IMODE, int JMODE, class WaitF>
__device__ __forceinline__ bool tile_mainloop(Acc& acc, const double* Ag, size_t lda, const double* Bg, size_t ldb,
                                              int nk16, double* smem, int* /*s_abort*/, WaitF /*waitf*/) {
    constexpr bool IS = (IMODE == STREAM_M);
    constexpr bool JS = (JMODE == STREAM_M || JMODE == STREAM_K);
    constexpr bool JK = (JMODE == STREAM_K || JMODE == RES_K);
    constexpr bool JRES = !JS;
    constexpr bool SS = IS && JS;                        // both operands streamed
    constexpr int NST = SS ? SS_STAGES : STAGES;         // stage buffers per operand
    constexpr int LA = NST / 2 - 1;                      // pairs issued ahead of the one being computed
    constexpr int SJ = (SS && !JK) ? STAGE_I : STAGE_J;  // doubles per j stage
    constexpr int PB = JS ? (JK ? PK : PM) : PM;
    static_assert((size_t)SS_STAGES * (STAGE_I + STAGE_J) * sizeof(double) <= TILE_SMEM_BYTES, "six stage pairs must fit the block");
    double* r0 = smem;
    double* r1 = smem + R0_DBL;
    double* istage = SS ? smem : (JRES ? r1 : r0);       // i stages move to region 1 when j is resident in region 0
    double* jstage = SS ? smem + SS_STAGES * STAGE_I : r1;
    const TileCoord tc;
    StageCopy<false> ci;
    StageCopy<JK> cj;
    if (IS) ci.init(Ag, lda);
    if (JS) cj.init(Bg, ldb);

    // stage index of k16 step kk is kk % NST; ld_s / cp_s walk it without a division
    int ld_s = 0, cp_s = 0;
    auto issue_pair = [&](bool real) {
        if (real) {
            if (IS) ci.issue_pair(istage + ld_s * STAGE_I, istage + (ld_s + 1) * STAGE_I);
            if (JS) cj.issue_pair(jstage + ld_s * SJ, jstage + (ld_s + 1) * SJ);
            ld_s = (ld_s + 2 == NST) ? 0 : ld_s + 2;
        }
        cp_async_commit();                               // an (empty) group per iteration keeps the accounting uniform
    };
    auto a_of = [&](int kk, int h) -> const double* {
        return IS ? (istage + (cp_s + h) * STAGE_I) : (Ag + (size_t)(kk + h) * KT * PM);
    };
    auto b_of = [&](int kk, int h) -> const double* {
        return JS ? (jstage + (cp_s + h) * SJ) : (JK ? (Bg + (kk + h) * KT) : (Bg + (size_t)(kk + h) * KT * PM));
    };

    __syncthreads();                                     // the previous user of the shared-memory block is done
#pragma unroll
    for (int p = 0; p < LA; ++p) issue_pair(2 * p < nk16);
    for (int kk = 0; kk < nk16; kk += 2) {
        cp_async_wait<LA - 1>();                         // pair kk has landed (for this thread) ...
        __syncthreads();                                 // ... and for everybody; pair kk-2 is consumed by all
        compute_ks<JK, 0, 2>(acc, a_of(kk, 0), PM, b_of(kk, 0), PB, tc);
        issue_pair(kk + 2 * LA < nk16);
        compute_ks<JK, 2, 4>(acc, a_of(kk, 0), PM, b_of(kk, 0), PB, tc);
        compute_ks<JK, 0, 4>(acc, a_of(kk, 1), PM, b_of(kk, 1), PB, tc);
        cp_s = (cp_s + 2 == NST) ? 0 : cp_s + 2;
    }
    cp_async_wait<0>();
    __syncthreads();
    return true;
}

struct NoWait {
    __device__ __forceinline__ bool operator()(int) const { return true; }
};

// ---- tile <-> memory helpers -------------------------------------------------------------------

// Store accumulators as a column-major tile: element (row, col) -> dst[col*ld + row].
// SIGN = +1 stores acc, -1 stores -acc.  Works for global or shared destinations.
template <bool JK, int SIGN>
__device__ __forceinline__ void store_tile(const Acc& acc, double* dst, size_t ld, const TileCoord& tc) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int c = tc.col<JK>(mt);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            double* d = dst + (size_t)c * ld + tc.row(p, 0);
            double2 lo, hi;
            lo.x = SIGN * GPR_ACC(acc, mt, p, 0); lo.y = SIGN * GPR_ACC(acc, mt, p, 1);
            hi.x = SIGN * GPR_ACC(acc, mt, p, 2); hi.y = SIGN * GPR_ACC(acc, mt, p, 3);
            reinterpret_cast<double2*>(d)[0] = lo;
            reinterpret_cast<double2*>(d)[1] = hi;
        }
    }
}

// dst_smem (column-major, pitch PM) = G - acc, where G is a column-major global tile (read through L2).
// Two passes: the accumulators are first parked in shared memory (negated), then the tile is read with
// coalesced 16-byte loads and added in place.  (Subtracting in registers, in fragment layout, makes ptxas
// hoist all 64 global loads next to the 128 accumulator registers: 255 registers and spilled copy
// strides that were reloaded inside every DMMA loop of the kernel.)  Ends with a __syncthreads().
template <bool JK>
__device__ __forceinline__ void residual_to_smem(const Acc& acc, const double* G, size_t ld, double* dst,
                                                 const TileCoord& tc) {
    store_tile<JK, -1>(acc, dst, PM, tc);
    __syncthreads();
    for (int idx = threadIdx.x; idx < TB * TB / 2; idx += NTHREADS) {
        const int r2 = idx & 63, c = idx >> 6;
        const double2 g = __ldcg(reinterpret_cast<const double2*>(G + (size_t)c * ld) + r2);
        double2* d = reinterpret_cast<double2*>(dst + c * PM) + r2;
        double2 v = *d;
        v.x += g.x; v.y += g.y;
        *d = v;
    }
    __syncthreads();
}

}  // namespace gpr

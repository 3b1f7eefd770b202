/* gpr_selftest.h — engine self-tests and pipe probes exported by libgpr_b200.so for tests/ and bench.py only.
 * Not part of the drop-in boundary (include/gpr_c_api.h): nothing in the reference binds to these. */
#ifndef GPR_SELFTEST_H
#define GPR_SELFTEST_H
#ifdef __cplusplus
extern "C" {
#endif

/* ---- self-tests of the tile engine (used by tests/, device pointers, one 128-tile granularity) -- */
int gpr_selftest_gemm(const double* hA, const double* hB, int b_kmajor, double* hC, int m_tiles, int n_tiles, int k);
int gpr_selftest_leaf(double* h_tile_inout, double* h_inv_out, int* info);
int gpr_selftest_factor(double* hA_inout, int n_tiles, double* h_linv_or_null, int serial, long long* pivot);
/* Timeline of the tile-task Cholesky: 4 ns stamps per task (n_tiles(n_tiles+1)/2 tasks, column-major
 * task order); leaf_cycles[2] = SM cycles of the in-CTA 128x128 Cholesky and triangular inverse. */
int gpr_selftest_factor_trace(int n_tiles, long long* h_trace, long long* leaf_cycles_or_null);
/* Raw pipe probes (CUDA-event timed): which = 0 FP64 tensor (DMMA.8x8x4), 1 FP64 FMA, 2 / 3 both pipes
 * mixed (16 DMMA with 32 / 128 DFMA per thread); total TFLOP/s. */
int gpr_selftest_peak(int which, int ctas_per_sm, double* tflops);

/* INT8 tensor-core engine (gpr_ozaki.cu: tcgen05.mma kind::i8 + TMEM + TMA): raw level accumulators
 * C[l] = sum_{t+u=l} A_t B_u^T for int8 slice tensors A [S][M][K], B [S][N][K] (K contiguous); hC: [levels][M][N].
 * skip_zero_blocks: build the nonzero-slice map of A and skip the MMAs of all-zero (row tile, k-block, slice) blocks. */
int gpr_selftest_i8gemm(const signed char* hA, const signed char* hB, int S, int levels, int M, int N, int K, int tri,
                        int skip_zero_blocks, int* hC);

#ifdef __cplusplus
}
#endif
#endif

"""GPU suite, part 1: the DMMA tile engine and the shared-memory leaves, piece by piece against numpy
(float64 reference of the same op).  Tolerances: products of O(1) numbers with k terms accumulate
k*eps relative error; factorisations are compared with LAPACK at cond*eps."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    return M @ M.T / n + np.eye(n)


@pytest.mark.parametrize("kmajor", [False, True])
@pytest.mark.parametrize("shape", [(128, 128, 32), (256, 384, 64), (128, 256, 160), (384, 128, 1056)])
def test_tile_gemm_matches_numpy(gpr, kmajor, shape):
    M, N, k = shape
    rng = np.random.default_rng(M + N + k)
    A, B = rng.standard_normal((M, k)), rng.standard_normal((N, k))
    C = gpr.selftest_gemm(A, B, kmajor)
    assert relerr(C, A @ B.T) <= 1e-14 * np.sqrt(k)


def test_tile_gemm_is_exact_on_small_integers(gpr):
    rng = np.random.default_rng(0)
    A = rng.integers(-8, 9, size=(256, 96)).astype(float)
    B = rng.integers(-8, 9, size=(128, 96)).astype(float)
    for kmajor in (False, True):
        assert np.array_equal(gpr.selftest_gemm(A, B, kmajor), A @ B.T)      # fragment maps are exact


def test_leaf_cholesky_and_inverse(gpr):
    A = _spd(128, 3)
    L, inv, info = gpr.selftest_leaf(A)
    Lr = np.linalg.cholesky(A)
    assert info == 0
    assert relerr(np.tril(L), Lr) <= 1e-14 and np.abs(np.triu(L, 1)).max() == 0.0
    assert relerr(inv, np.linalg.inv(Lr)) <= 1e-13 and np.abs(np.triu(inv, 1)).max() == 0.0


@pytest.mark.parametrize("col", [0, 15, 16, 77, 127])
def test_leaf_reports_first_bad_pivot(gpr, col):
    A = _spd(128, 4)
    A[col, col] = -5.0
    assert gpr.selftest_leaf(A)[2] == col + 1


@pytest.mark.parametrize("nb,serial", [(1, True), (2, True), (3, True), (3, False), (8, False), (24, False)])
def test_tile_cholesky_and_inverse(gpr, nb, serial):
    """The persistent tile-task kernels (dependency flags) and the one-launch-per-column debug mode."""
    A = _spd(128 * nb, 10 + nb)
    L, X, piv = gpr.selftest_factor(A, True, serial)
    Lr = np.linalg.cholesky(A)
    assert piv == 0
    assert relerr(L, Lr) <= 1e-13
    assert relerr(np.tril(X), np.linalg.inv(Lr)) <= 1e-12
    assert relerr(np.tril(X) @ L, np.eye(len(A))) <= 1e-12


def test_tile_cholesky_is_bit_reproducible(gpr):
    A = _spd(128 * 12, 99)
    L1, _, _ = gpr.selftest_factor(A, False, False)
    L2, _, _ = gpr.selftest_factor(A, False, False)
    L3, _, _ = gpr.selftest_factor(A, False, True)
    assert np.array_equal(L1, L2) and np.array_equal(L1, L3)     # scheduling never changes the arithmetic


def _kernel_matrix(n, seed, nugget):
    """Gaussian-kernel matrix of random points: smooth kernel + small nugget = the ill-conditioned kind of matrix a GP fit meets."""
    rng = np.random.default_rng(seed)
    P = rng.random((n, 3))
    D2 = ((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)
    return np.exp(-D2 / 0.5) + nugget * np.eye(n)


@pytest.mark.parametrize("nb,panel,S,nugget", [(12, 4, 7, 1e-6), (12, 5, 7, 1e-3), (24, 8, 7, 1e-5), (16, 3, 8, 1e-6), (8, 8, 7, 1e-4)])
def test_int8_assisted_cholesky_is_as_accurate_as_fp64(gpr, nb, panel, S, nugget):
    """launch_cholesky_int8: the flops left of each panel of tile columns run on the INT8 tensor cores (base-254 digit slices of
    L, exact integer products, FP64 recombination), the panels themselves on the FP64 tile kernel.  Its backward error
    ||L L^T - A|| must match the all-FP64 factorisation's on ill-conditioned kernel matrices; one panel = the FP64 kernel itself."""
    n = 128 * nb
    A = _kernel_matrix(n, 100 + nb, nugget)
    L64, _, p64 = gpr.selftest_factor(A, False, False)
    L8, _, p8 = gpr.selftest_factor(A, False, 100 * panel + S)
    assert p64 == 0 and p8 == 0
    if panel >= nb:
        assert np.array_equal(L8, L64)
        return
    An = np.linalg.norm(A)
    b64 = np.linalg.norm(L64 @ L64.T - A) / An
    b8 = np.linalg.norm(L8 @ L8.T - A) / An
    print(nb, panel, S, "backward error fp64 %.3e int8-assisted %.3e" % (b64, b8), "factor difference %.3e" % relerr(L8, L64))
    assert b8 <= 2.0 * b64 + 2e-16
    assert relerr(L8, L64) <= 1e-6                    # both are within cond(A) * eps of the exact factor
    L8b, _, _ = gpr.selftest_factor(A, False, 100 * panel + S)
    assert np.array_equal(L8, L8b)                    # integer sums + one producer per tile: bit-reproducible


def test_int8_assisted_cholesky_keeps_rows_of_very_different_size_accurate(gpr):
    """Per-point noise / very different diagonal entries: every row of L is sliced against its own power-of-two scale
    (above sqrt(K_ii)), so a row 10^6 times smaller than the largest keeps all its digits.  Measured on A = D K D with
    D = diag(10^u), u in [-3, 3]: the error of L L^T, entry (i, j) relative to d_i d_j, matches the all-FP64 factorisation's
    (one common scale would lose log2(d_max / d_i) bits in row i)."""
    n = 128 * 10
    rng = np.random.default_rng(7)
    K = _kernel_matrix(n, 21, 1e-4)
    d = 10.0 ** rng.uniform(-3.0, 3.0, n)
    A = K * d[:, None] * d[None, :]
    L64, _, p64 = gpr.selftest_factor(A, False, False)
    L8, _, p8 = gpr.selftest_factor(A, False, 100 * 3 + 7)
    assert p64 == 0 and p8 == 0
    # the 48 smallest rows behind the first panel, L L^T evaluated in extended precision (a float64 product would drown the signal)
    rows = 384 + np.argsort(d[384:])[:48]
    Al = A.astype(np.longdouble)

    def scaled(L):
        Ll = L.astype(np.longdouble)
        # columns behind the first panel only: entries (i, j < 384) involve nothing but the first panel, which is all-FP64
        return float(np.abs((Ll[rows] @ Ll[384:].T - Al[rows, 384:]) / (d[rows, None] * d[None, 384:])).max())
    e64, e8 = scaled(L64), scaled(L8)
    print("scaled backward error of the smallest rows: fp64 %.3e, int8-assisted %.3e" % (e64, e8))
    assert e8 <= 3.0 * e64 + 1e-16            # one common scale: ~1e-11 here


def test_int8_assisted_cholesky_reports_a_bad_pivot_in_a_later_panel(gpr):
    A = _spd(128 * 8, 5)
    A[700, 700] = -1.0
    _, _, piv = gpr.selftest_factor(A, False, 100 * 2 + 7)
    assert piv == 701


def test_tile_cholesky_rejects_indefinite(gpr):
    A = _spd(128 * 4, 5)
    A[300, 300] = -1.0
    _, _, piv = gpr.selftest_factor(A, False, False)
    assert piv == 301


def test_fp64_pipe_probes(gpr):
    dmma, dfma = gpr.selftest_peak(0, 4), gpr.selftest_peak(1, 4)
    assert 10.0 < dmma < 80.0 and 10.0 < dfma < 80.0       # B200: ~37 / ~34 TF/s; wide bounds: a power-capped box must not fail a parity suite


@pytest.mark.parametrize("S,M,N,K,tri", [(1, 128, 64, 64, False), (2, 256, 128, 512, False), (3, 384, 192, 384, True),
                                         (7, 512, 320, 1024, True), (8, 256, 64, 2048, False)])
def test_int8_tensor_core_engine_is_exact(gpr, S, M, N, K, tri):
    """gpr_ozaki.cu: tcgen05.mma kind::i8 into TMEM int32 accumulators, operands by TMA (SWIZZLE_64B, K-major), one
    accumulator per level l = t + u.  Integer arithmetic: the result must equal numpy's int64 product exactly."""
    rng = np.random.default_rng(S * 1000 + K)
    A = rng.integers(-64, 65, size=(S, M, K), dtype=np.int8)
    B = rng.integers(-64, 65, size=(S, N, K), dtype=np.int8)
    if S >= 3:
        # digit slices that vanish in whole (128-row, 64-k) blocks, like the leading digits of L^-1 far from the diagonal:
        # their MMAs are skipped (except in the first k-block of a task, which initialises the accumulators)
        for (t, r, kb) in [(0, 0, 1), (0, 1, 0), (1, 1, 3), (0, M // 128 - 1, K // 64 - 1), (2, 0, 0)]:
            A[t, r * 128:(r + 1) * 128, kb * 64:(kb + 1) * 64] = 0
        A[0, 128:256, 128:] = 0
    C = gpr.selftest_i8gemm(A, B, S, tri, skip_zero_blocks=S >= 3)
    A64 = A.astype(np.int64)
    if tri:
        for r in range(M // 128):
            A64[:, r * 128:(r + 1) * 128, 128 * (r + 1):] = 0
    for l in range(S):
        E = sum(A64[t] @ B[l - t].astype(np.int64).T for t in range(l + 1))
        assert np.array_equal(C[l].astype(np.int64), E), l


@pytest.mark.parametrize("S,M,N,K,tri,kchunk", [(2, 256, 128, 1024, False, 4), (6, 512, 192, 1024, True, 2),
                                                (7, 384, 64, 1536, True, 6), (8, 256, 320, 2048, False, 10)])
def test_int8_engine_with_chunked_accumulation_is_exact(gpr, monkeypatch, S, M, N, K, tri, kchunk):
    """Rows longer than the int32 accumulators allow (k > 22016 at 6 slices of |d| <= 127) run the k-chunked kernel: the
    accumulators are drained every `kchunk` k-blocks.  GPR_OZ_KCHUNK forces that kernel with a short chunk, so that tasks
    span several chunks (and, for `tri`, a ragged last chunk); the summed level accumulators must equal the int64 product."""
    monkeypatch.setenv("GPR_OZ_KCHUNK", str(kchunk))
    rng = np.random.default_rng(S * 77 + K)
    A = rng.integers(-127, 128, size=(S, M, K), dtype=np.int8)
    B = rng.integers(-127, 128, size=(S, N, K), dtype=np.int8)
    # zero blocks, among them the first block of a chunk (which must not be skipped: it initialises the accumulators)
    for (t, r, kb) in [(0, 0, kchunk), (0, 1, 0), (1, 1, kchunk + 1), (0, M // 128 - 1, K // 64 - 1)]:
        A[t, r * 128:(r + 1) * 128, kb * 64:(kb + 1) * 64] = 0
    C = gpr.selftest_i8gemm(A, B, S, tri, skip_zero_blocks=True)
    A64 = A.astype(np.int64)
    if tri:
        for r in range(M // 128):
            A64[:, r * 128:(r + 1) * 128, 128 * (r + 1):] = 0
    for l in range(S):
        E = sum(A64[t] @ B[l - t].astype(np.int64).T for t in range(l + 1))
        assert np.array_equal(C[l].astype(np.int64), E), l

"""CPU suite, part 1: pins the oracle (oracle/gpr_oracle.cpp) against
  * the fixtures produced by the reference's own header (tests/golden/ref_*.npz, oracle/make_golden.py),
  * that header itself when oracle/_ref is present, and
  * the closed-form known answers of SURVEY Appendix B.
Tolerances are written next to each check."""
import math

import numpy as np
import pytest

from conftest import load_golden, relerr

CASES = ["ref_mugD_thinplate", "ref_kettle_gaussian", "ref_jug_gaussian", "ref_jug_laplace"]


def _fit(orc, g, **kw):
    P = g["P"]
    return orc.Oracle(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"], g["kind"], g["p0"], g["p1"], **kw)


NODE_CASES = ["ref_mugD_thinplate_R2_node", "ref_jug_thinplate_R2_node"]      # indefinite K


@pytest.mark.parametrize("case", CASES + NODE_CASES)
def test_oracle_reproduces_reference_fixture_exactly(orc, case):
    """Expansion-form distance + pivoted LDLT + no FMA contraction = the reference's arithmetic:
    every output must agree to the last bits (1e-13 relative leaves room for libm's exp only)."""
    g = load_golden(case)
    normals = "normals" in g
    o = _fit(orc, g, factor="ldlt", dist="expansion", with_normals=normals)
    Q = g["Q"]
    f, v, grad = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    got = o.get(K=True)
    assert abs(o.R - g["R"]) == 0.0
    assert relerr(got["K"][0], g["K_row0"]) <= 1e-15
    assert relerr(np.diag(got["K"]), g["K_diag"]) <= 1e-15
    assert relerr(got["alpha"], g["alpha"]) <= 1e-13
    assert relerr(f, g["f"]) <= 1e-13 and relerr(f, g["f_mean_only"]) <= 1e-13
    assert np.abs(v - g["v"]).max() <= 1e-13 * max(1.0, np.abs(g["v"]).max())
    assert relerr(grad, g["grad"]) <= 1e-13
    N, Tx, Ty = orc.tangent_basis(grad)
    assert np.abs(Tx - g["Tx"]).max() <= 1e-12 and np.abs(Ty - g["Ty"]).max() <= 1e-12
    if normals:
        assert np.abs(got["normals"] - g["normals"]).max() <= 1e-12


@pytest.mark.parametrize("case", CASES)
def test_difference_form_and_llt_agree_with_reference_to_conditioning(orc, case):
    """Deviation (i) (difference-form distance) and LLT instead of pivoted LDLT move alpha by
    O(cond(K)*eps) only: <= 1e-10 relative at these sizes (cond ~ 1e5); mean 1e-11."""
    g = load_golden(case)
    Q = g["Q"]
    for factor in ("ldlt", "llt"):
        o = _fit(orc, g, factor=factor, dist="diff")
        assert o.info == 0
        assert relerr(o.alpha, g["alpha"]) <= 1e-10
        f, v, grad = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
        assert relerr(f, g["f"]) <= 1e-11
        assert np.abs(v - g["v"]).max() <= 1e-9 * np.abs(g["v"]).max()
        assert relerr(grad, g["grad"]) <= 1e-11
        assert int((np.sign(f) != np.sign(g["f"]))[np.abs(g["f"]) > 1e-9].sum()) == 0


def test_update_matches_reference_fixture(orc):
    """update<>() = append + refactorise (gp_regressor.hpp:442-459); R is not refreshed (:454-455)."""
    g = load_golden("ref_mugD_thinplate")
    o = _fit(orc, g, factor="ldlt", dist="expansion")
    Pu = g["Pu"]
    o.update(Pu[:, 0], Pu[:, 1], Pu[:, 2], g["yu"], g["su"])
    assert o.n == len(g["P"]) + len(Pu)
    assert relerr(o.alpha, g["alpha_updated"]) <= 1e-13
    assert o.R == g["R_updated"] == g["R"]
    Q = g["Q"]
    f, _, _ = o.predict(Q[:, 0], Q[:, 1], Q[:, 2])
    assert relerr(f, g["f_updated"]) <= 1e-13


def test_live_reference_header_when_present(orc):
    """Where oracle/_ref was built (this container), run the reference's header directly on a fresh
    random input, all four evaluate overloads, and compare bit for bit."""
    if not orc.have_reference():
        pytest.skip("oracle/_ref/libgpr_ref.so not built")
    rng = np.random.default_rng(5)
    P = rng.uniform(-1, 1, size=(150, 3))
    y = rng.integers(0, 2, size=150).astype(float)
    s2 = np.full(150, 0.1)
    Q = rng.uniform(-1, 1, size=(40, 3))
    for kind, p0, p1 in (("thin_plate", 4.0, 0.0), ("gaussian", 0.8, 1.3), ("laplace", 1.2, 0.7)):
        ref = orc.Reference(kind, p0, p1).fit(P[:, 0], P[:, 1], P[:, 2], y, s2, with_normals=True)
        o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, kind, p0, p1, factor="ldlt", dist="expansion", with_normals=True)
        assert relerr(o.alpha, ref.get()["alpha"]) == 0.0
        f, v, grad = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
        assert relerr(f, ref.evaluate(Q[:, 0], Q[:, 1], Q[:, 2], 1)[0]) == 0.0
        fr, vr, gr, tx, ty = ref.evaluate(Q[:, 0], Q[:, 1], Q[:, 2], 4)
        assert relerr(f, fr) == 0.0 and np.abs(v - vr).max() == 0.0 and relerr(grad, gr) == 0.0
        N, Tx, Ty = orc.tangent_basis(grad)
        assert np.abs(Tx - tx).max() <= 1e-15 and np.abs(Ty - ty).max() <= 1e-15


def test_reference_error_messages_fixture(orc):
    z = np.load(__import__("os").path.join(__import__("conftest").GOLD, "ref_error_messages.npz"))
    assert list(z["messages"]) == ["Empty data pointer", "All input data is empty!", "Query is already labeled!",
                                   "Empty Model pointer"]


# ---- closed forms -------------------------------------------------------------
def test_kernel_known_values(orc):
    k = orc.kernel_value
    assert k("thin_plate", 2.0, 0, 0.0) == 8.0 and k("thin_plate", 2.0, 0, 2.0) == 0.0        # B.1
    assert k("thin_plate", 2.0, 0, 1.0) == 4.0 and k("thin_plate", 2.0, 0, 1.0, True) == -6.0
    assert k("thin_plate", 2.0, 0, 0.0, True) == -12.0 and k("thin_plate", 2.0, 0, 2.0, True) == 0.0
    assert k("gaussian", 1.5, 2.0, 0.0) == 2.25                                               # B.2
    assert abs(k("gaussian", 1.5, 2.0, 4.0) - 2.25 / math.e) <= 1e-15
    assert k("laplace", 1.5, 2.0, 0.0) == 3.0 and abs(k("laplace", 1.5, 2.0, 2.0) - 3.0 / math.e) <= 1e-15
    assert abs(k("gaussian", 1.5, 2.0, 1.0, True) + k("gaussian", 1.5, 2.0, 1.0) / 4.0) <= 1e-16
    assert abs(k("laplace", 1.5, 2.0, 1.0, True) + k("laplace", 1.5, 2.0, 1.0) / 2.0) <= 1e-16


@pytest.mark.parametrize("factor", ["ldlt", "llt"])
def test_single_point_posterior(orc, factor):
    """B.3: ThinPlate R=2, y=1, s=0.1, d=1: alpha=1/8.1, f=4/8.1, v=8-16/8.1, G=-(6/8.1)(q-p)."""
    o = orc.Oracle([0.5], [0.0], [0.0], [1.0], [0.1], "thin_plate", 2.0, 0.0, factor=factor)
    f, v, g = o.predict([1.5], [0.0], [0.0], var=True, grad=True)
    assert abs(o.alpha[0] - 1 / 8.1) <= 1e-16
    assert abs(f[0] - 4 / 8.1) <= 1e-15 and abs(v[0] - (8 - 16 / 8.1)) <= 1e-14
    assert np.abs(g[0] - np.array([-6 / 8.1, 0, 0])).max() <= 1e-15


def test_two_point_posterior(orc):
    """B.4: K=[[a,b],[b,a]] => alpha = (a*y - b*Jy)/(a^2-b^2)."""
    d0, s, R = 0.7, 0.05, 2.0
    a = R ** 3 + s
    b = 2 * d0 ** 3 - 3 * R * d0 ** 2 + R ** 3
    y = np.array([1.0, -0.5])
    o = orc.Oracle([0.0, d0], [0, 0], [0, 0], y, [s, s], "thin_plate", R, 0.0)
    want = (a * y - b * y[::-1]) / (a * a - b * b)
    assert relerr(o.alpha, want) <= 1e-14


def test_interpolation_without_noise(orc):
    """B.5: sigma2 empty => f(p_i)=y_i and v(p_i)=0 up to cond*eps."""
    rng = np.random.default_rng(3)
    P = rng.uniform(-1, 1, size=(60, 3))
    y = rng.standard_normal(60)
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, None, "gaussian", 1.0, 1.0, factor="llt")
    f, v, _ = o.predict(P[:, 0], P[:, 1], P[:, 2], var=True)
    assert np.abs(f - y).max() <= 1e-9 and np.abs(v).max() <= 1e-9


def test_tangent_basis_known_answers(orc):
    """B.6."""
    N, Tx, Ty = orc.tangent_basis(np.array([[0, 0, 2.0], [3.0, 0, 0], [0.3, -1.2, 0.8]]))
    assert np.allclose(N[0], [0, 0, 1]) and np.allclose(Tx[0], [1, 0, 0]) and np.allclose(Ty[0], [0, 1, 0])
    assert np.allclose(N[1], [1, 0, 0]) and np.allclose(Tx[1], [0, 1, 0]) and np.allclose(Ty[1], [0, 0, 1])
    for a, b in ((N, Tx), (N, Ty), (Tx, Ty)):
        assert abs(np.dot(a[2], b[2])) <= 1e-15
    for a in (N, Tx, Ty):
        assert abs(np.linalg.norm(a[2]) - 1) <= 1e-15


def test_spd_guard_node_setting(orc):
    """B.7 / F2: the node's ThinPlate(2.0) on mugD + r=2 sphere is indefinite: LLT stops at the first
    external point (row 263), while R = max distance factorises."""
    g = load_golden("ref_mugD_thinplate")
    P = g["P"]
    bad = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"], "thin_plate", 2.0, 0.0, factor="llt")
    assert bad.info == 263
    good = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"], "thin_plate", g["R"], 0.0, factor="llt")
    assert good.info == 0


def test_long_double_attribution(orc):
    """F9: both factorisations sit within cond*eps of an 80-bit solve; used by the GPU parity tests."""
    P, y, s2 = __import__("gpr_b200").workloads.synthetic_cloud(512, seed=0)
    R = __import__("gpr_b200").workloads.SYNTH_R
    ld = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", R, 0.0, factor="llt", precision="longdouble").alpha
    for factor in ("ldlt", "llt"):
        a = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", R, 0.0, factor=factor).alpha
        assert relerr(a, ld) <= 1e-10


def test_blas_flavour_matches(orc):
    """The numpy/scipy (OpenBLAS) flavour used as the timed CPU baseline computes the same numbers."""
    g = load_golden("ref_kettle_gaussian")
    m = orc.blas_fit(g["P"], g["y"], g["s2"], g["kind"], g["p0"], g["p1"])
    f, v = orc.blas_predict(m, g["Q"])
    assert relerr(m["alpha"], g["alpha"]) <= 1e-10 and relerr(f, g["f"]) <= 1e-11
    assert np.abs(v - g["v"]).max() <= 1e-9 * np.abs(g["v"]).max()


@pytest.mark.parametrize("case", NODE_CASES)
def test_node_configuration_is_indefinite_and_block_elimination_matches(orc, case):
    """The node's ThinPlate(2.0) + external sphere (src/gp_node.cpp:919, :821-849): K has negative eigenvalues, the
    reference's pivoted LDLT solves it anyway.  The block elimination used on the GPU (leading SPD block by Cholesky,
    the trailing 15 points as one dense pivot block, csrc/gpr_tail.cu) is restated here in numpy and must reproduce
    the reference's alpha / mean / variance to cond*eps."""
    g = load_golden(case)
    P, y, s2, Q = g["P"], g["y"], g["s2"], g["Q"]
    R = g["p0"]
    kern = lambda d: 2 * d ** 3 - 3 * R * d ** 2 + R ** 3
    dist = lambda A, B: np.sqrt(((A[:, None, :] - B[None, :, :]) ** 2).sum(-1))
    K = kern(dist(P, P)) + np.diag(s2)
    assert (np.linalg.eigvalsh(K) < 0).sum() == 3
    p = len(P) - 15
    L = np.linalg.cholesky(K[:p, :p])                       # the leading block (object points) is SPD
    X = np.linalg.inv(L)
    B = X @ K[:p, p:]
    S = K[p:, p:] - B.T @ B
    assert (np.linalg.eigvalsh(S) < 0).sum() == 3           # all the indefiniteness sits in the Schur complement
    Z = X.T @ B
    Sinv = np.linalg.inv(S)
    a2 = Sinv @ (y[p:] - B.T @ (X @ y[:p]))
    a1 = X.T @ (X @ y[:p]) - Z @ a2
    alpha = np.concatenate([a1, a2])
    assert relerr(alpha, g["alpha"]) <= 1e-10
    Kq = kern(dist(Q, P))
    f = Kq @ alpha
    w = Kq[:, p:] - Kq[:, :p] @ Z
    v = R ** 3 - ((Kq[:, :p] @ X.T) ** 2).sum(1) - np.einsum("qa,ab,qb->q", w, Sinv, w)
    assert relerr(f, g["f"]) <= 1e-10 and np.abs(v - g["v"]).max() <= 1e-9 * np.abs(g["v"]).max()

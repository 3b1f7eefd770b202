"""GPU suite, part 2: parity of the CUDA path with the oracle and with the reference fixtures, always
through the C-ABI (ctypes binding) — the tests read like the reference's own usage:
create<>() then the evaluate overloads, then update<>().

Tolerances (BASELINE.json north_star): relative error <= 1e-9 on mean and alpha, <= 1e-7 on variance,
identical sign of f on every query.  alpha is compared with max(1e-9, 4 x the oracle's own distance to an
80-bit solve) where cond(K) makes 1e-9 unreachable for ANY double implementation."""
import os
import subprocess
import threading

import numpy as np
import pytest

from conftest import ROOT, load_golden, relerr

pytestmark = pytest.mark.gpu

TOL_ALPHA, TOL_MEAN, TOL_VAR = 1e-9, 1e-9, 1e-7
CASES = ["ref_mugD_thinplate", "ref_kettle_gaussian", "ref_jug_gaussian", "ref_jug_laplace"]


@pytest.fixture(scope="module")
def ctx(gpr):
    return gpr.Context()


def _reg(gpr, ctx, g):
    return gpr.GPRegressor(g["kind"], g["p0"], g["p1"], ctx=ctx)


def _signs_agree(f, fo):
    mask = np.abs(fo) > 1e-9
    return int((np.sign(f) != np.sign(fo))[mask].sum()) == 0


@pytest.mark.parametrize("case", CASES)
def test_parity_with_reference_fixtures(gpr, orc, ctx, case):
    """Configs 1 and 2: PCD clouds, node preprocessing, all four evaluate overloads + normals."""
    g = load_golden(case)
    P, Q = g["P"], g["Q"]
    normals = "normals" in g
    reg = _reg(gpr, ctx, g)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"], with_normals=normals)
    got = m.get()
    assert abs(got["R"] - g["R"]) <= 1e-15 * g["R"]
    assert relerr(got["alpha"], g["alpha"]) <= TOL_ALPHA
    f1 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2])
    f2, v2 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    f3, v3, g3 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    f4, v4, g4, tx, ty = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, tangent=True)
    for f in (f1, f2, f3, f4):
        assert relerr(f, g["f"]) <= TOL_MEAN and _signs_agree(f, g["f"])
    for v in (v2, v3, v4):
        assert np.abs(v - g["v"]).max() <= TOL_VAR * np.abs(g["v"]).max()
    for gr in (g3, g4):
        assert relerr(gr, g["grad"]) <= TOL_MEAN
    assert np.abs(tx - g["Tx"]).max() <= 1e-8 and np.abs(ty - g["Ty"]).max() <= 1e-8
    if normals:
        assert np.abs(got["normals"] - g["normals"]).max() <= 1e-8
    # the thin-plate covariance build is bit-identical to the oracle's difference-form K
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"], g["kind"], g["p0"], g["p1"], factor="llt", dist="diff")
    Lo = np.tril(o.get(factor=True)["factor"])
    assert relerr(m.factor(), Lo) <= 1e-11
    assert relerr(got["alpha"], o.alpha) <= TOL_ALPHA


def test_single_query_calls_match_batched(gpr, ctx):
    """Every real caller passes q = 1 (src/gp_node.cpp:1074): warp-per-query mean + matrix-vector variance."""
    g = load_golden("ref_mugD_thinplate")
    P, Q = g["P"], g["Q"]
    reg = _reg(gpr, ctx, g)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"])
    for i in (0, 17, 100, 251):
        f, v, gr = reg.evaluate(m, Q[i:i + 1, 0], Q[i:i + 1, 1], Q[i:i + 1, 2], var=True, grad=True)
        assert abs(f[0] - g["f"][i]) <= TOL_MEAN * np.abs(g["f"]).max()
        assert abs(v[0] - g["v"][i]) <= TOL_VAR * np.abs(g["v"]).max()
        assert np.abs(gr[0] - g["grad"][i]).max() <= TOL_MEAN * np.abs(g["grad"]).max()
    for q in (2, 3, 5, 8, 9, 130):
        f, v = reg.evaluate(m, Q[:q, 0], Q[:q, 1], Q[:q, 2], var=True)
        assert relerr(f, g["f"][:q]) <= TOL_MEAN and np.abs(v - g["v"][:q]).max() <= TOL_VAR * np.abs(g["v"]).max()
    # the fused q <= 8 kernel: every overload, against the reference fixture (mean-only, +grad, +tangent basis)
    for q in (1, 4, 7):
        s = slice(40, 40 + q)
        f1 = reg.evaluate(m, Q[s, 0], Q[s, 1], Q[s, 2])
        f4, v4, g4, tx, ty = reg.evaluate(m, Q[s, 0], Q[s, 1], Q[s, 2], var=True, tangent=True)
        scale = np.abs(g["f"]).max()
        assert np.abs(f1 - g["f"][s]).max() <= TOL_MEAN * scale and np.abs(f4 - g["f"][s]).max() <= TOL_MEAN * scale
        assert np.abs(v4 - g["v"][s]).max() <= TOL_VAR * np.abs(g["v"]).max()
        assert np.abs(g4 - g["grad"][s]).max() <= TOL_MEAN * np.abs(g["grad"]).max()
        assert np.abs(tx - g["Tx"][s]).max() <= 1e-8 and np.abs(ty - g["Ty"][s]).max() <= 1e-8


def test_single_query_path_at_scale_and_after_append(gpr, orc, ctx):
    """q <= 8 on a model with several k-splits and row blocks (n = 3000), before and after an incremental
    append that grows the capacity (the fused kernel's scratch and ticket must survive the growth)."""
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(3100, seed=9)
    rng = np.random.default_rng(9)
    perm = rng.permutation(len(P)); P, y, s2 = P[perm], y[perm], s2[perm]
    n0 = 3000
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m = reg.create(P[:n0, 0], P[:n0, 1], P[:n0, 2], y[:n0], s2[:n0])
    Q = W.grid_slab(10, 2, 3)
    for stage in range(2):
        n = m.n
        o = orc.Oracle(P[:n, 0], P[:n, 1], P[:n, 2], y[:n], s2[:n], "thin_plate", W.SYNTH_R, 0.0, factor="llt")
        fo, vo, go = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True, threads=4)
        for a, q in ((0, 1), (1, 2), (3, 8), (11, 5)):
            f, v, gr = reg.evaluate(m, Q[a:a + q, 0], Q[a:a + q, 1], Q[a:a + q, 2], var=True, grad=True)
            assert np.abs(f - fo[a:a + q]).max() <= TOL_MEAN * np.abs(fo).max()
            assert np.abs(v - vo[a:a + q]).max() <= TOL_VAR * np.abs(vo).max()
            assert np.abs(gr - go[a:a + q]).max() <= TOL_MEAN * np.abs(go).max()
        if stage == 0:
            reg.update(m, P[n0:n0 + 90, 0], P[n0:n0 + 90, 1], P[n0:n0 + 90, 2], y[n0:n0 + 90], s2[n0:n0 + 90])
            assert m.n == n0 + 90


def test_update_matches_reference_fixture(gpr, ctx):
    g = load_golden("ref_mugD_thinplate")
    P, Q, Pu = g["P"], g["Q"], g["Pu"]
    reg = _reg(gpr, ctx, g)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"])
    reg.update(m, Pu[:, 0], Pu[:, 1], Pu[:, 2], g["yu"], g["su"])
    assert m.n == len(P) + len(Pu)
    assert relerr(m.alpha, g["alpha_updated"]) <= TOL_ALPHA
    assert m.R == pytest.approx(g["R"], rel=1e-15)                 # not refreshed (gp_regressor.hpp:454-455)
    assert relerr(reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2]), g["f_updated"]) <= TOL_MEAN


def test_not_spd_is_reported_never_nan(gpr, ctx, monkeypatch):
    """Appendix B.7: the node's ThinPlate(2.0) setting is indefinite; pivot = first external point.  With the
    trailing-block elimination switched off (GPR_NO_TAIL=1) that is an error, never NaN; an indefinite matrix
    with more than 256 offending points is an error in any case."""
    g = load_golden("ref_mugD_thinplate")
    P = g["P"]
    reg = gpr.GPRegressor("thin_plate", 2.0, ctx=ctx)
    monkeypatch.setenv("GPR_NO_TAIL", "1")
    with pytest.raises(gpr.GPRegressionException) as e:
        reg.create(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"])
    assert e.value.code == gpr.GPR_ERR_NOT_SPD and e.value.pivot == 263
    monkeypatch.delenv("GPR_NO_TAIL")
    W = gpr.workloads
    Ps, ys, ss = W.synthetic_cloud(1024, seed=3)
    Ps, ys, ss = Ps[:768], ys[:768], ss[:768]                        # the unit-sphere part (scaled below: diameter 0.9 <= R)
    rng = np.random.default_rng(3)
    d = rng.standard_normal((300, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    far = np.vstack([d * 40.0, Ps * 0.45])                           # 300 mutually distant outliers FIRST, then a compact cloud
    reg2 = gpr.GPRegressor("thin_plate", 2.0, ctx=ctx)
    with pytest.raises(gpr.GPRegressionException) as e:              # more than 256 offending points: an error in any case
        reg2.create(far[:, 0], far[:, 1], far[:, 2], np.concatenate([np.ones(300), ys]), np.full(len(far), 0.1))
    assert e.value.code == gpr.GPR_ERR_NOT_SPD and 1 <= e.value.pivot <= len(far)


@pytest.mark.parametrize("order", ["externals_first", "shuffled"])
def test_indefinite_matrix_with_offending_points_anywhere(gpr, ctx, order):
    """The node's indefinite setting with the training set re-ordered: the 15 external points first, or all 277
    points shuffled.  The offending points are moved to the end of the INTERNAL order (the caller never sees it:
    alpha comes back in the caller's order) and eliminated as the trailing pivot block; results must equal the
    reference's pivoted-LDLT outputs for the same set."""
    g = load_golden("ref_mugD_thinplate_R2_node")
    P, Q, n = g["P"], g["Q"], len(g["P"])
    idx = np.concatenate([np.arange(n - 15, n), np.arange(n - 15)]) if order == "externals_first" else np.random.default_rng(4).permutation(n)
    reg = _reg(gpr, ctx, g)
    m = reg.create(P[idx, 0], P[idx, 1], P[idx, 2], g["y"][idx], g["s2"][idx], with_normals=True)
    assert m.n_tail == 15 and m.n == n
    got = m.get()
    assert relerr(got["alpha"], g["alpha"][idx]) <= TOL_ALPHA
    assert np.abs(got["normals"] - g["normals"][idx]).max() <= 1e-8
    f, v, gr = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    assert relerr(f, g["f"]) <= TOL_MEAN and _signs_agree(f, g["f"])
    assert np.abs(v - g["v"]).max() <= TOL_VAR * np.abs(g["v"]).max() and relerr(gr, g["grad"]) <= TOL_MEAN
    f1, v1 = reg.evaluate(m, Q[:1, 0], Q[:1, 1], Q[:1, 2], var=True)
    assert abs(f1[0] - g["f"][0]) <= TOL_MEAN * np.abs(g["f"]).max() and abs(v1[0] - g["v"][0]) <= TOL_VAR * np.abs(g["v"]).max()


@pytest.mark.parametrize("case", ["ref_mugD_thinplate_R2_node", "ref_jug_thinplate_R2_node"])
def test_node_configuration_indefinite_matrix(gpr, ctx, case):
    """the ROS node's REAL setting — ThinPlate(2.0), 15 external points at r = 2 — gives an
    indefinite K (3 negative eigenvalues).  The reference's pivoted LDLT handles it; here the 15 trailing points
    are eliminated as one dense pivot block (block L D L^T, csrc/gpr_tail.cu).  Parity against the reference's
    own outputs for that setting: all four evaluate overloads, training normals, q = 1 calls."""
    g = load_golden(case)
    P, Q = g["P"], g["Q"]
    reg = _reg(gpr, ctx, g)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"], with_normals=True)
    assert m.n_tail == 15 and m.n == len(P)
    got = m.get()
    assert abs(got["R"] - g["R"]) <= 1e-15 * g["R"]
    assert relerr(got["alpha"], g["alpha"]) <= TOL_ALPHA
    assert np.abs(got["normals"] - g["normals"]).max() <= 1e-8
    f1 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2])
    f4, v4, g4, tx, ty = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, tangent=True)
    for f in (f1, f4):
        assert relerr(f, g["f"]) <= TOL_MEAN and _signs_agree(f, g["f"])
    assert np.abs(v4 - g["v"]).max() <= TOL_VAR * np.abs(g["v"]).max()
    assert relerr(g4, g["grad"]) <= TOL_MEAN
    assert np.abs(tx - g["Tx"]).max() <= 1e-8 and np.abs(ty - g["Ty"]).max() <= 1e-8
    for i in (0, 5, 100):                                        # the node's call pattern: one query per call
        f, v = reg.evaluate(m, Q[i:i + 1, 0], Q[i:i + 1, 1], Q[i:i + 1, 2], var=True)
        assert abs(f[0] - g["f"][i]) <= TOL_MEAN * np.abs(g["f"]).max()
        assert abs(v[0] - g["v"][i]) <= TOL_VAR * np.abs(g["v"]).max()
    # a big batch (thread-per-query + tile variance path) agrees with the small-batch path
    big = np.vstack([Q] * 40)
    fb, vb = reg.evaluate(m, big[:, 0], big[:, 1], big[:, 2], var=True)
    assert relerr(fb[:len(Q)], g["f"]) <= TOL_MEAN and np.abs(vb[:len(Q)] - g["v"]).max() <= TOL_VAR * np.abs(g["v"]).max()
    assert np.array_equal(fb[:len(Q)], fb[-len(Q):]) and np.array_equal(vb[:len(Q)], vb[-len(Q):])
    # update() on such a model is incremental too: the two new surface points come AFTER the external points in the caller's
    # order, yet only the 15 external points stay in the trailing block (they are moved behind the new points
    # in the internal order); alpha comes back in the caller's order and agrees with a fresh fit of all 279
    reg.update(m, [0.3, 0.0], [0.1, 0.5], [-0.2, 0.4], [0.0, 0.0], [0.05, 0.05])
    assert m.n == len(P) + 2 and m.n_tail == 15
    P2 = np.vstack([P, [[0.3, 0.1, -0.2], [0.0, 0.5, 0.4]]])
    fresh = reg.create(P2[:, 0], P2[:, 1], P2[:, 2], np.concatenate([g["y"], [0.0, 0.0]]), np.concatenate([g["s2"], [0.05, 0.05]]))
    assert fresh.n_tail == 15 and relerr(m.alpha, fresh.alpha) <= 1e-12


def test_closed_form_posteriors(gpr, ctx):
    """Appendix B.3 / B.4 on the GPU."""
    reg = gpr.GPRegressor("thin_plate", 2.0, ctx=ctx)
    m = reg.create([0.5], [0.0], [0.0], [1.0], [0.1])
    f, v, gr = reg.evaluate(m, [1.5], [0.0], [0.0], var=True, grad=True)
    assert abs(m.alpha[0] - 1 / 8.1) <= 1e-15 and abs(f[0] - 4 / 8.1) <= 1e-14
    assert abs(v[0] - (8 - 16 / 8.1)) <= 1e-13 and np.abs(gr[0] - [-6 / 8.1, 0, 0]).max() <= 1e-14
    d0, s, R = 0.7, 0.05, 2.0
    a, b = R ** 3 + s, 2 * d0 ** 3 - 3 * R * d0 ** 2 + R ** 3
    y = np.array([1.0, -0.5])
    m2 = reg.create([0.0, d0], [0, 0], [0, 0], y, [s, s])
    assert relerr(m2.alpha, (a * y - b * y[::-1]) / (a * a - b * b)) <= 1e-13


def test_interpolation_without_noise(gpr, ctx):
    """Appendix B.5 (sigma2 empty): f(p_i) = y_i, v(p_i) = 0."""
    rng = np.random.default_rng(3)
    P = rng.uniform(-1, 1, size=(60, 3))
    y = rng.standard_normal(60)
    reg = gpr.GPRegressor("gaussian", 1.0, 1.0, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, None)
    f, v = reg.evaluate(m, P[:, 0], P[:, 1], P[:, 2], var=True)
    assert np.abs(f - y).max() <= 1e-9 and np.abs(v).max() <= 1e-9


@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 300, 1025])
def test_ragged_sizes(gpr, orc, ctx, n):
    """Training sets that do not fill whole 128-tiles (identity padding) and ragged query counts."""
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(max(n, 4), seed=n)
    P, y, s2 = P[-n:], y[-n:], s2[-n:]
    Q = W.grid_slab(7, 0, 7)[: 3 * n + 5]
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2, with_normals=True)
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", W.SYNTH_R, 0.0, factor="llt", with_normals=True)
    assert relerr(m.alpha, o.alpha) <= TOL_ALPHA and abs(m.R - o.R) <= 1e-15
    f, v, gr = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    fo, vo, go = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True, threads=8)
    assert relerr(f, fo) <= TOL_MEAN and np.abs(v - vo).max() <= TOL_VAR * np.abs(vo).max() and relerr(gr, go) <= TOL_MEAN
    if n > 1:
        assert np.abs(m.get()["normals"] - o.get()["normals"]).max() <= 1e-8


def test_synthetic_2048_all_outputs(gpr, orc, ctx):
    """Config-3 generator at a size the oracle finishes in seconds; alpha judged against the 80-bit solve."""
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(2048, seed=0)
    Q = W.grid_slab(24, 0, 24)[::5]
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", W.SYNTH_R, 0.0, factor="llt")
    ld = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", W.SYNTH_R, 0.0, factor="llt", precision="longdouble")
    err_ref, err_gpu = relerr(o.alpha, ld.alpha), relerr(m.alpha, ld.alpha)
    assert err_gpu <= max(TOL_ALPHA, 4 * err_ref)
    f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    fo, vo, _ = ld.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, threads=8)
    assert relerr(f, fo) <= TOL_MEAN and np.abs(v - vo).max() <= TOL_VAR * np.abs(vo).max() and _signs_agree(f, fo)
    assert v.min() > 0.0


def test_query_batches_and_shards_are_bit_identical(gpr):
    """each query is computed by one device with identical code, so splitting the query set
    (across batches, calls or GPUs) never changes a bit — for a given model state: the form of the variance (INT8 tensor
    cores / FP64 product / forward substitution) is chosen on the first large call and is sticky afterwards.
    Shards stay >= 9472 queries (thread-per-query path)."""
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(1536, seed=2)
    Q = W.grid_slab(40, 0, 16)                      # 25,600 queries
    os.environ["GPR_QUERY_TILE"] = "4096"           # force several variance batches per call
    try:
        small = gpr.Context()
    finally:
        del os.environ["GPR_QUERY_TILE"]
    big = gpr.Context()
    out = []
    for c in (small, big):
        reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=c)
        m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
        out.append(reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True))
        if c is big:
            h = len(Q) // 2
            parts = [reg.evaluate(m, Q[a:b, 0], Q[a:b, 1], Q[a:b, 2], var=True, grad=True) for a, b in ((0, h), (h, len(Q)))]
            out.append(tuple(np.concatenate([p[i] for p in parts]) for i in range(3)))
    for other in out[1:]:
        for a, b in zip(out[0], other):
            assert np.array_equal(a, b)


def test_fit_is_bit_reproducible(gpr, ctx):
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(3000, seed=4)
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    a1 = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2).alpha
    a2 = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2).alpha
    assert np.array_equal(a1, a2)


def test_concurrent_single_query_calls(gpr, ctx):
    """The node evaluates from hundreds of threads on one shared model (src/gp_node.cpp:1027-1038)."""
    g = load_golden("ref_jug_gaussian")
    P, Q = g["P"], g["Q"]
    reg = _reg(gpr, ctx, g)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"])
    reg.prepare_variance(m)
    res = [None] * len(Q)

    def work(i):
        res[i] = reg.evaluate(m, Q[i:i + 1, 0], Q[i:i + 1, 1], Q[i:i + 1, 2], var=True)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(Q))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    f = np.array([r[0][0] for r in res]); v = np.array([r[1][0] for r in res])
    assert relerr(f, g["f"]) <= TOL_MEAN and np.abs(v - g["v"]).max() <= TOL_VAR * np.abs(g["v"]).max()


def _lattice_queries(W, res, count, seed):
    """`count` points of the bench lattice (res^3 on [-1.2, 1.2]^3): half from slabs through the object, where the
    surface f = 0 is crossed, half anywhere."""
    rng = np.random.default_rng(seed)
    out = []
    for z in rng.integers(res // 4, 3 * res // 4, size=4):
        slab = W.grid_slab(res, int(z), int(z) + 1)
        r = np.linalg.norm(slab, axis=1)
        near = np.flatnonzero(np.abs(r - 1.0) < 0.05)
        out.append(slab[rng.choice(near, size=count // 8, replace=False)])
        out.append(slab[rng.choice(len(slab), size=count // 8, replace=False)])
    Q = np.vstack(out)
    extra = count - len(Q)
    if extra > 0:
        Q = np.vstack([Q, W.grid_slab(res, 1, 2)[:extra]])
    return Q[:count]


def _headline_parity(gpr, orc, ctx, monkeypatch, n, res, nq, with_blas):
    """Parity at a headline size against an INDEPENDENT extended-precision check (oracle.certify): the GPU only
    supplies approximate solves w ~ K^-1 k*, the CPU forms the residuals k* - K w with ~106-bit sums over a K it
    assembles itself and returns f and v exact to second order in the residual.  Every production variance path
    (forward substitution over L, product with L^-1, the fused single-query kernel) is then held to the north-star
    tolerances against that: 1e-9 on the mean, 1e-7 on the variance, identical sign."""
    W = gpr.workloads
    R = W.SYNTH_R
    P, y, s2 = W.synthetic_cloud(n, seed=0)
    reg = gpr.GPRegressor("thin_plate", R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    alpha = m.alpha
    Q = _lattice_queries(W, res, nq, seed=n)
    from oracle import oracle as O
    Ks = O._kern("thin_plate", R, 0.0, O._pdist(P, Q))
    Wsol = reg.solve(m, Ks)                                         # two triangular solves per column on the GPU
    cert = orc.certify(P, y, s2, "thin_plate", R, 0.0, Q, alpha, Wsol)
    fc, vc = cert["f"], cert["v"]
    # the certificate itself: neglected second-order terms far below the tolerances
    assert cert["v_second_order"] <= 1e-12 * np.abs(vc).max() and cert["f_second_order"] <= 1e-12 * np.abs(fc).max()
    assert vc.min() > 0.0 and vc.max() < R ** 3
    report = {"n": n, "queries": nq, "resid_inf": cert["resid_inf"], "v_second_order": cert["v_second_order"]}
    # alpha: one refinement step with the extended-precision residual r = y - K alpha gives alpha's error to first order
    dalpha = reg.solve(m, cert["R"][:, -1])
    cond_eps = 40.0 * n / float(np.min(s2)) * 2.2e-16              # lambda_max ~ 40 n, lambda_min >= sigma2
    report["alpha_rel_err"] = float(np.abs(dalpha).max() / np.abs(alpha).max())
    assert report["alpha_rel_err"] <= max(TOL_ALPHA, 50 * cond_eps)
    assert m.state().linv is None                                    # nothing has built L^-1 so far
    monkeypatch.setenv("GPR_VAR_MODE", "trsm")
    f_t, v_t = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    assert ctx.timings()["linv_ms"] == 0.0 or m.state().linv is None
    monkeypatch.setenv("GPR_VAR_MODE", "product")
    f_p, v_p = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    assert m.state().linv is not None
    monkeypatch.setenv("GPR_VAR_MODE", "ozaki")                       # INT8 tensor cores (tcgen05), FP64-equivalent by slicing
    f_o, v_o = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    assert ctx.timings()["ozaki_ms"] > 0.0
    monkeypatch.delenv("GPR_VAR_MODE")
    f_1 = np.zeros(8); v_1 = np.zeros(8)
    for i in range(8):                                               # the callers' pattern: one query per call
        fi, vi = reg.evaluate(m, Q[i:i + 1, 0], Q[i:i + 1, 1], Q[i:i + 1, 2], var=True)
        f_1[i], v_1[i] = fi[0], vi[0]
    for name, f, v, sl in (("trsm", f_t, v_t, slice(None)), ("product", f_p, v_p, slice(None)), ("ozaki_int8", f_o, v_o, slice(None)),
                           ("single", f_1, v_1, slice(0, 8))):
        report["mean_rel_" + name] = float(np.abs(f - fc[sl]).max() / np.abs(fc).max())
        report["var_rel_" + name] = float(np.abs(v - vc[sl]).max() / np.abs(vc).max())
        assert report["mean_rel_" + name] <= TOL_MEAN, report
        assert report["var_rel_" + name] <= TOL_VAR, report
        assert _signs_agree(f, fc[sl])
    if with_blas:
        # the second anchor: OpenBLAS dpotrf / dpotrs / dtrsm on the host cores (SURVEY §8d "best-effort CPU")
        mb = orc.blas_fit(P, y, s2, "thin_plate", R, 0.0)
        fb, vb = orc.blas_predict(mb, Q, var=True)
        report["alpha_rel_vs_dpotrs"] = relerr(alpha, mb["alpha"])
        report["mean_rel_vs_blas"] = relerr(f_t, fb)
        report["var_rel_vs_blas"] = float(np.abs(v_t - vb).max() / np.abs(vb).max())
        assert report["alpha_rel_vs_dpotrs"] <= max(TOL_ALPHA, 50 * cond_eps)
        assert report["mean_rel_vs_blas"] <= TOL_MEAN and report["var_rel_vs_blas"] <= TOL_VAR and _signs_agree(f_t, fb)
    print("headline parity:", report)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        import json
        with open(os.path.join(out, "parity_full_size_n%d.json" % n), "w") as fh:
            json.dump(report, fh)
    return reg, m, P, y, s2


def test_headline_parity_config3(gpr, orc, ctx, monkeypatch):
    """BASELINE config 3 at full size (n = 16,384, ThinPlate, queries of the 256^3 lattice): alpha, mean, variance and
    sign against the extended-precision certificate AND against OpenBLAS dpotrf/dpotrs/dtrsm; plus the size-independent
    properties (variances positive and below k(0); the mean ~0 on the unit sphere, ~1 on the outer one)."""
    reg, m, P, y, s2 = _headline_parity(gpr, orc, ctx, monkeypatch, 16384, 256, 255, with_blas=True)
    W = gpr.workloads
    Q = W.grid_slab(256, 128, 129)[:148 * 128]
    f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)    # a full batch: by default the INT8 tensor-core form
    t = ctx.timings()
    assert t["predict_var_ms"] > 0 and t["ozaki_ms"] > 0
    assert np.isfinite(f).all() and v.min() > 0.0 and v.max() < W.SYNTH_R ** 3
    monkeypatch.setenv("GPR_VAR_MODE", "product")
    f2, v2 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    monkeypatch.delenv("GPR_VAR_MODE")
    assert ctx.timings()["ozaki_ms"] == 0.0 and np.array_equal(f, f2)
    assert np.abs(v - v2).max() <= 1e-8 * np.abs(v2).max()         # 10x inside the tolerance on all 18,944 queries
    fs = reg.evaluate(m, P[::64, 0], P[::64, 1], P[::64, 2])
    assert np.abs(fs - y[::64]).max() < 0.2          # sigma2 = 0.1 smoothing, not interpolation
    m.close()


def test_headline_parity_config5(gpr, orc, ctx, monkeypatch):
    """BASELINE config 5 at full size (n = 65,536: K = 32 GiB factorised in place; + 32 GiB once L^-1 is built for the
    product form): the same certificate on 79 lattice queries of the 512^3 grid (a host dpotrf would take ~15 min)."""
    import torch
    free, total = torch.cuda.mem_get_info(0)
    if free < 90 * (1 << 30):
        pytest.skip("needs ~80 GB of free device memory")
    reg, m, P, y, s2 = _headline_parity(gpr, orc, ctx, monkeypatch, 65536, 512, 79, with_blas=False)
    m.close()


@pytest.mark.parametrize("case", ["ref_mugD_thinplate", "ref_mugD_thinplate_R2_node"])
def test_cxx_dropin_headers_end_to_end(gpr, orc, tmp_path, case):
    """The reference-facing C++ API (include/gp_regression/*.h) compiled as C++11 and run on the GPU — on the SPD
    configuration and on the node's own indefinite ThinPlate(2.0) setting (where a plain Cholesky drop-in would throw)."""
    from test_host import _build_driver
    exe = _build_driver(str(tmp_path))
    g = load_golden(case)
    has_upd = "Pu" in g
    P, Q = g["P"], g["Q"][:40]
    Pu = g["Pu"] if has_upd else np.zeros((0, 3))
    with open(tmp_path / "in.txt", "w") as fh:
        row = lambda *v: " ".join(repr(float(x)) for x in v) + "\n"
        fh.write("0 %r 0.0\n%d %d %d 1\n" % (float(g["p0"]), len(P), len(Q), len(Pu)))
        for p, l, s in zip(P, g["y"], g["s2"]):
            fh.write(row(p[0], p[1], p[2], l, s))
        for p in Q:
            fh.write(row(*p))
        if has_upd:
            for p, l, s in zip(Pu, g["yu"], g["su"]):
                fh.write(row(p[0], p[1], p[2], l, s))
    subprocess.run([exe, str(tmp_path / "in.txt"), str(tmp_path / "out.txt")], check=True)
    text = open(tmp_path / "out.txt").read()
    assert "exception" not in text, text[:500]
    rows = {ln.split()[0]: np.array(ln.split()[1:], dtype=float) for ln in text.splitlines()}
    q = len(Q)
    assert abs(rows["R"][0] - g["R"]) <= 1e-14 * g["R"] and relerr(rows["alpha"], g["alpha"]) <= TOL_ALPHA
    assert np.abs(rows["normals"].reshape(3, -1).T - g["normals"]).max() <= 1e-8
    for k in ("f1", "f2", "f3", "f4"):
        assert relerr(rows[k], g["f"][:q]) <= TOL_MEAN
    for k in ("v2", "v3", "v4"):
        assert np.abs(rows[k] - g["v"][:q]).max() <= TOL_VAR * np.abs(g["v"]).max()
    assert relerr(rows["N3"].reshape(3, -1).T, g["grad"][:q]) <= TOL_MEAN
    assert np.abs(rows["Tx"].reshape(3, -1).T - g["Tx"][:q]).max() <= 1e-8
    assert abs(rows["single"][0] - g["f"][0]) <= 1e-9 and abs(rows["single"][1] - g["v"][0]) <= 1e-7 * np.abs(g["v"]).max()
    if has_upd:
        assert relerr(rows["alpha_updated"], g["alpha_updated"]) <= TOL_ALPHA
        assert relerr(rows["f_updated"], g["f_updated"][:q]) <= TOL_MEAN and rows["R_updated"][0] == rows["R"][0]
    # the batched sampler through the C++ header agrees with the Python mirror of the same C-ABI call
    reg = gpr.GPRegressor("thin_plate", g["p0"])
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"])
    pts, fs, vs = reg.sample_isosurface(m)
    assert int(rows["iso"][0]) == len(pts) and abs(rows["iso"][1] - vs.sum()) <= 1e-9 * max(1.0, abs(vs.sum()))


# ---- K5: incremental append (rank-k row append of L and L^-1) vs refit, SURVEY row a13 / config 4 -------------
def _append_case(gpr, n=900, seed=5):
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(n + 200, seed=seed)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(P))
    P, y, s2 = P[perm], y[perm], s2[perm]
    return W, P, y, s2


@pytest.mark.parametrize("batches", [(32,), (7, 1, 50), (32, 32, 32, 32, 32)])
def test_incremental_append_matches_refit_and_oracle(gpr, orc, ctx, batches, monkeypatch):
    """n = 900 base (tile padding 124) so that the appended rows cross a 128-tile boundary and grow the
    capacity; slabs of 32 and ragged slabs.  The incremental model must agree with a from-scratch fit of the
    same points (alpha to cond*eps, factor to 1e-11) and with the oracle's update() (a refit, like the
    reference, gp_regressor.hpp:442-459), and later variance queries must use the appended L^-1 rows."""
    W, P, y, s2 = _append_case(gpr)
    n0 = 900
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m = reg.create(P[:n0, 0], P[:n0, 1], P[:n0, 2], y[:n0], s2[:n0])
    o = orc.Oracle(P[:n0, 0], P[:n0, 1], P[:n0, 2], y[:n0], s2[:n0], "thin_plate", W.SYNTH_R, 0.0, factor="llt")
    a = n0
    for k in batches:
        reg.update(m, P[a:a + k, 0], P[a:a + k, 1], P[a:a + k, 2], y[a:a + k], s2[a:a + k])
        o.update(P[a:a + k, 0], P[a:a + k, 1], P[a:a + k, 2], y[a:a + k], s2[a:a + k])
        a += k
        assert ctx.timings()["append_ms"] > 0.0            # the incremental path ran (a refit leaves it at 0)
    assert m.n == a
    fresh = reg.create(P[:a, 0], P[:a, 1], P[:a, 2], y[:a], s2[:a])
    assert relerr(m.factor(), fresh.factor()) <= 1e-11
    assert relerr(m.alpha, fresh.alpha) <= TOL_ALPHA and relerr(m.alpha, o.alpha) <= TOL_ALPHA
    Q = W.grid_slab(12, 0, 12)[::5]
    f, v, g = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    f2, v2, g2 = reg.evaluate(fresh, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    fo, vo, go = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True, threads=4)
    assert relerr(f, f2) <= TOL_MEAN and relerr(f, fo) <= TOL_MEAN and relerr(g, go) <= TOL_MEAN
    assert np.abs(v - v2).max() <= TOL_VAR * np.abs(vo).max() and np.abs(v - vo).max() <= TOL_VAR * np.abs(vo).max()
    f1, v1 = reg.evaluate(m, Q[:3, 0], Q[:3, 1], Q[:3, 2], var=True)          # q <= 8: matrix-vector path on L^-1
    assert np.abs(v1 - vo[:3]).max() <= TOL_VAR * np.abs(vo).max()


def test_incremental_append_equals_forced_refit(gpr, ctx, monkeypatch):
    W, P, y, s2 = _append_case(gpr, seed=6)
    n0, k = 700, 40
    reg = gpr.GPRegressor("gaussian", 1.0, 1.0, ctx=ctx)
    sl = lambda a, b: (P[a:b, 0], P[a:b, 1], P[a:b, 2], y[a:b], s2[a:b])
    m1 = reg.create(*sl(0, n0))
    reg.update(m1, *sl(n0, n0 + k))
    monkeypatch.setenv("GPR_APPEND_REFIT", "1")
    m2 = reg.create(*sl(0, n0))
    reg.update(m2, *sl(n0, n0 + k))
    assert relerr(m1.alpha, m2.alpha) <= TOL_ALPHA and relerr(m1.factor(), m2.factor()) <= 1e-11


def test_incremental_append_is_bit_reproducible_and_reserve(gpr, ctx):
    W, P, y, s2 = _append_case(gpr, seed=7)
    n0 = 640
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    outs = []
    for rep in range(2):
        m = reg.create(P[:n0, 0], P[:n0, 1], P[:n0, 2], y[:n0], s2[:n0])
        if rep == 1:
            reg.reserve(m, 2048)                       # pre-allocated capacity must not change a single bit
        for a in range(n0, n0 + 96, 32):
            reg.update(m, P[a:a + 32, 0], P[a:a + 32, 1], P[a:a + 32, 2], y[a:a + 32], s2[a:a + 32])
        Q = W.grid_slab(8, 0, 8)
        f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        outs.append((m.alpha.copy(), m.factor(), f, v))
    for x0, x1 in zip(*outs):
        assert np.array_equal(x0, x1)


def test_incremental_append_failure_leaves_model_intact(gpr, ctx):
    """A point farther than R from the cloud makes the thin-plate matrix indefinite: NOT_SPD with
    the global pivot index — found in the SECOND slab, after the first was already committed on the device —
    and the model answers exactly as before the call."""
    W, P, y, s2 = _append_case(gpr, seed=8)
    n0 = 500
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m = reg.create(P[:n0, 0], P[:n0, 1], P[:n0, 2], y[:n0], s2[:n0])
    Q = W.grid_slab(6, 0, 6)
    before = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    alpha0, L0 = m.alpha.copy(), m.factor()
    bad = np.vstack([P[n0:n0 + 41], [[3 * W.SYNTH_R, 0.0, 0.0]], P[n0 + 41:n0 + 45]])
    with pytest.raises(gpr.GPRegressionException) as e:
        reg.update(m, bad[:, 0], bad[:, 1], bad[:, 2], np.zeros(len(bad)), np.full(len(bad), 0.1))
    assert e.value.code == gpr.GPR_ERR_NOT_SPD and e.value.pivot == n0 + 42
    assert m.n == n0 and np.array_equal(m.alpha, alpha0) and np.array_equal(m.factor(), L0)
    after = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
    reg.update(m, P[n0:n0 + 16, 0], P[n0:n0 + 16, 1], P[n0:n0 + 16, 2], y[n0:n0 + 16], s2[n0:n0 + 16])   # still usable
    fresh = reg.create(P[:n0 + 16, 0], P[:n0 + 16, 1], P[:n0 + 16, 2], y[:n0 + 16], s2[:n0 + 16])
    assert m.n == n0 + 16 and relerr(m.alpha, fresh.alpha) <= TOL_ALPHA


@pytest.mark.parametrize("tag", ["spd", "R2_node"])
def test_batched_isosurface_sampler_against_reference_and_oracle(gpr, orc, ctx, tag):
    """the node's fakeDeterministicSampling (src/gp_node.cpp:998-1100) as ONE call — same lattice
    (x outermost, accumulated axis values), same |f| <= 0.01 criterion, variance as the kept points' intensity.
    Checked (a) against the reference's own evaluate() over the node's 29^3 lattice on mugD (fixture generated by
    oracle/make_golden.py: lattice_cases) and (b) against the CPU oracle evaluated on the whole lattice here —
    for the SPD setting and for the node's own ThinPlate(2.0) (indefinite K, pivoted LDLT in the oracle)."""
    g = load_golden("ref_mugD_lattice_" + tag)
    P, y, s2, R = g["P"], g["y"], g["s2"], float(g["p0"])
    W = gpr.workloads
    reg = gpr.GPRegressor("thin_plate", R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    assert (m.n_tail > 0) == (tag == "R2_node")
    Q = W.node_grid()                                         # 29^3, same loop nest
    assert len(Q) == int(g["lattice"])
    pts, fs, vs = reg.sample_isosurface(m)
    # (a) the reference
    idx, fr, vr = g["idx"], g["f"], g["v"]
    keep_r = np.abs(fr) <= 0.01
    edge_r = np.abs(np.abs(fr) - 0.01) < 1e-9                 # points sitting on the threshold may flip
    assert int(keep_r.sum()) == int(g["kept"])
    if not edge_r.any():
        assert np.array_equal(pts, Q[idx[keep_r]])
        assert np.abs(fs - fr[keep_r]).max() <= TOL_MEAN * np.abs(fr).max()
        assert np.abs(vs - vr[keep_r]).max() <= TOL_VAR * np.abs(vr).max()
    # (b) the oracle over the whole lattice
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", R, 0.0, factor="ldlt")
    fo, vo, _ = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, threads=os.cpu_count() or 1)
    keep = np.abs(fo) <= 0.01
    edge = np.abs(np.abs(fo) - 0.01) < 1e-9
    assert 0 < keep.sum() < len(Q) // 4
    assert (keep & ~edge).sum() <= len(pts) <= (keep | edge).sum()
    if not edge.any():
        assert np.array_equal(pts, Q[keep])
        assert np.abs(fs - fo[keep]).max() <= TOL_MEAN * np.abs(fo).max()
        assert np.abs(vs - vo[keep]).max() <= TOL_VAR * np.abs(vo).max()
    pts2, fs2, _ = reg.sample_isosurface(m, var=False, capacity=5)      # truncated output, mean only
    assert len(pts2) == 5 and np.array_equal(pts2, pts[:5]) and np.array_equal(fs2, fs[:5])


def test_batched_isosurface_sampler_fine_lattice(gpr, orc, ctx):
    """A finer lattice on a bigger SPD model (several device chunks), survivors checked against the oracle."""
    W = gpr.workloads
    Ps, ys, ss = W.synthetic_cloud(1500, seed=11)
    reg2 = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m2 = reg2.create(Ps[:, 0], Ps[:, 1], Ps[:, 2], ys, ss)
    pts3, fs3, vs3 = reg2.sample_isosurface(m2, lo=-1.2, hi=1.2, step=2.4 / 149, tol=0.002)
    assert len(pts3) > 100 and np.abs(fs3).max() <= 0.002 and vs3.min() > 0.0
    r = np.linalg.norm(pts3, axis=1)
    assert 0.8 < r.min() and r.max() < 1.2                    # the zero level set hugs the unit sphere of the cloud
    o = orc.Oracle(Ps[:, 0], Ps[:, 1], Ps[:, 2], ys, ss, "thin_plate", W.SYNTH_R, 0.0, factor="llt")
    sub = slice(0, len(pts3), max(1, len(pts3) // 400))
    fo, vo, _ = o.predict(pts3[sub, 0], pts3[sub, 1], pts3[sub, 2], var=True, threads=os.cpu_count() or 1)
    assert np.abs(fs3[sub] - fo).max() <= 1e-9 and np.abs(vs3[sub] - vo).max() <= TOL_VAR * np.abs(vo).max()


def _oracle_project(o, p, g, f_tol, improve_tol, max_iter, step_mul):
    """include/atlas/atlas.hpp:201-276 transcribed, its two evaluate(q = 1) calls per iteration answered by the CPU
    oracle (mean only at :225; mean + variance + gradient at :259)."""
    cur, g = np.array(p, dtype=np.float64), np.array(g, dtype=np.float64)
    for it in range(max_iter):
        f_cur = o.predict(cur[:1], cur[1:2], cur[2:3])[0][0]
        if abs(f_cur) < f_tol:
            return cur, it + 1
        step = step_mul * f_cur * g
        if 1e-6 < np.linalg.norm(step) <= 100.0:
            cur = cur - step
        f_new, _, N = o.predict(cur[:1], cur[1:2], cur[2:3], var=True, grad=True)
        if 1e-5 < np.linalg.norm(N[0]) <= 100.0:
            g = N[0]
        if abs(f_new[0] - f_cur) < improve_tol:
            return cur, it + 1
    return cur, -max_iter


def test_batched_projection_matches_the_atlas_loop(gpr, orc, ctx):
    """AtlasBase::project for many points in one launch, against the reference's loop transcribed
    over single-query evaluations of the CPU ORACLE (same update rule, same three stopping criteria)."""
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(1200, seed=12)
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", W.SYNTH_R, 0.0, factor="llt")
    rng = np.random.default_rng(12)
    d = rng.standard_normal((6, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    start = d * rng.uniform(0.85, 1.25, size=(6, 1))                   # inside and outside the unit-sphere surface
    _, _, g0 = o.predict(start[:, 0], start[:, 1], start[:, 2], var=True, grad=True)
    kw = dict(f_tol=1e-3, improve_tol=1e-9, max_iter=80, step_mul=0.2)
    out, st = reg.project(m, start, g0, **kw)
    for i in range(len(start)):
        ref_out, ref_it = _oracle_project(o, start[i], g0[i], **kw)
        assert st[i] == ref_it
        assert np.abs(out[i] - ref_out).max() <= 1e-9
    f_end = o.predict(out[:, 0], out[:, 1], out[:, 2])[0]
    assert (st > 0).all() and np.abs(f_end).max() < 1e-3               # all converged onto the surface
    assert np.abs(np.linalg.norm(out, axis=1) - 1.0).max() < 0.1
    # budget exhaustion is reported, not hidden (the reference prints and returns the last iterate)
    out2, st2 = reg.project(m, start[:2], g0[:2], f_tol=1e-12, improve_tol=0.0, max_iter=5, step_mul=0.2)
    assert (st2 == -5).all()


@pytest.mark.parametrize("n,q", [(640, 1500), (1000, 300), (2304, 5000)])
def test_variance_by_forward_substitution_matches_oracle_and_product_form(gpr, orc, ctx, monkeypatch, n, q):
    """K3' without the explicit inverse (var_trsm_kernel: V = L^-1 K*^T by blocked forward substitution in the K* panel;
    reference: cholesker.solve(Kpq), gp_regressor.hpp:263, :316) against the oracle, and against the product with L^-1.
    Ragged n (padded tiles), ragged q (padded query tiles), more query tiles than SMs."""
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(n, seed=21)
    rng = np.random.default_rng(n)
    Q = rng.uniform(-1.2, 1.2, size=(q, 3))
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    monkeypatch.setenv("GPR_VAR_MODE", "trsm")
    f_t, v_t, g_t = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    assert m.state().linv is None                                   # the inverse factor was never built
    f_t2, v_t2 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    assert np.array_equal(v_t, v_t2)                                # bit-reproducible
    monkeypatch.setenv("GPR_VAR_MODE", "product")
    f_p, v_p = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    assert m.state().linv is not None
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", W.SYNTH_R, 0.0, factor="llt")
    sub = slice(0, q, max(1, q // 500))
    fo, vo, go = o.predict(Q[sub, 0], Q[sub, 1], Q[sub, 2], var=True, grad=True, threads=os.cpu_count() or 1)
    assert relerr(f_t[sub], fo) <= TOL_MEAN and relerr(g_t[sub], go) <= TOL_MEAN
    assert np.abs(v_t[sub] - vo).max() <= TOL_VAR * np.abs(vo).max()
    assert np.abs(v_t - v_p).max() <= 1e-10 * np.abs(v_p).max() and np.array_equal(f_t, f_p)
    # Gaussian kernel through the same path
    regg = gpr.GPRegressor("gaussian", 1.0, 1.0, ctx=ctx)
    mg = regg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    monkeypatch.setenv("GPR_VAR_MODE", "trsm")
    fg, vg = regg.evaluate(mg, Q[sub, 0], Q[sub, 1], Q[sub, 2], var=True)
    og = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "gaussian", 1.0, 1.0, factor="llt")
    fgo, vgo, _ = og.predict(Q[sub, 0], Q[sub, 1], Q[sub, 2], var=True, threads=os.cpu_count() or 1)
    assert relerr(fg, fgo) <= TOL_MEAN and np.abs(vg - vgo).max() <= TOL_VAR * np.abs(vgo).max()


@pytest.mark.parametrize("n,q,kind", [(1500, 3000, "thin_plate"), (2304, 20000, "thin_plate"), (1000, 700, "gaussian")])
def test_variance_on_int8_tensor_cores_matches_oracle(gpr, orc, ctx, monkeypatch, n, q, kind):
    """K3'' (gpr_ozaki.cu): the variance product on the INT8 tensor cores (tcgen05.mma kind::i8, TMEM accumulators, TMA
    operand feeds), FP64-equivalent by Ozaki slicing, against the oracle and the FP64 product form; ragged n and q; slice
    counts 7 (default) and 8; bit-reproducible."""
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(n, seed=31)
    Q = np.random.default_rng(n).uniform(-1.2, 1.2, size=(q, 3))
    p0, p1 = (W.SYNTH_R, 0.0) if kind == "thin_plate" else (1.0, 1.0)
    reg = gpr.GPRegressor(kind, p0, p1, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    monkeypatch.setenv("GPR_VAR_MODE", "product")
    f_p, v_p = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    monkeypatch.setenv("GPR_VAR_MODE", "ozaki")
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, kind, p0, p1, factor="llt")
    sub = slice(0, q, max(1, q // 400))
    fo, vo, _ = o.predict(Q[sub, 0], Q[sub, 1], Q[sub, 2], var=True, threads=os.cpu_count() or 1)
    for slices in (None, "7", "8"):
        if slices:
            monkeypatch.setenv("GPR_OZAKI_SLICES", slices)
        f_o, v_o = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        assert ctx.timings()["ozaki_ms"] > 0.0 and ctx.timings()["ozaki_slices"] == float(slices or 6)
        f_o2, v_o2 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        assert np.array_equal(v_o, v_o2) and np.array_equal(f_o, f_p)
        assert np.abs(v_o - v_p).max() <= 1e-8 * np.abs(v_p).max()
        assert relerr(f_o[sub], fo) <= TOL_MEAN and np.abs(v_o[sub] - vo).max() <= TOL_VAR * np.abs(vo).max()
    if kind == "thin_plate":
        # an incremental update changes X: the slices are cut again for the grown model
        monkeypatch.delenv("GPR_OZAKI_SLICES")
        Pn, yn, sn = W.touch_batches(1, 24)[0]
        reg.update(m, Pn[:, 0], Pn[:, 1], Pn[:, 2], yn, sn)
        o.update(Pn[:, 0], Pn[:, 1], Pn[:, 2], yn, sn)
        f_u, v_u = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        assert ctx.timings()["ozaki_ms"] > 0.0
        fo2, vo2, _ = o.predict(Q[sub, 0], Q[sub, 1], Q[sub, 2], var=True, threads=os.cpu_count() or 1)
        assert relerr(f_u[sub], fo2) <= TOL_MEAN and np.abs(v_u[sub] - vo2).max() <= TOL_VAR * np.abs(vo2).max()
        monkeypatch.setenv("GPR_VAR_MODE", "product")
        f_p, v_p = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        monkeypatch.setenv("GPR_VAR_MODE", "ozaki")
    # too few slices: the per-call FP64 spot check catches it and the call (and the model from then on) falls back
    monkeypatch.setenv("GPR_OZAKI_SLICES", "3")
    f_b, v_b = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    assert np.abs(v_b - v_p).max() <= 1e-10 * np.abs(v_p).max()     # the FP64 product form answered


def test_default_variance_path_needs_no_inverse_for_large_batches(gpr, ctx):
    """Without any override: a large batch on a fresh model takes the forward substitution (no L^-1 is built: time to
    first variance = the fit, one n x n matrix per model); a single-query call then builds L^-1 for the fused kernel,
    and later batches use the product form.  All three agree."""
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(900, seed=22)
    Q = W.grid_slab(48, 20, 22)                                       # 4608 queries >= GPR_TRSM_MIN_Q
    reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    f0, v0 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    assert m.state().linv is None
    f1, v1 = reg.evaluate(m, Q[:1, 0], Q[:1, 1], Q[:1, 2], var=True)
    assert m.state().linv is not None
    f2, v2 = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    assert np.abs(v0 - v2).max() <= 1e-10 * np.abs(v2).max() and abs(v1[0] - v0[0]) <= 1e-10 * np.abs(v2).max()
    assert np.array_equal(f0, f2)


def test_thread_fanout_of_single_query_calls_is_combined(gpr, ctx):
    """The reference's callers: hundreds of concurrent threads with ONE query each on one shared model
    (src/gp_node.cpp:1027-1038).  The micro-batcher inside gpr_predict combines what queues up behind a launch;
    every caller must get the value a batched call returns (to rounding: a lone request takes the fused q <= 8 kernel,
    combined ones the batched path), mixed overloads included."""
    g = load_golden("ref_mugD_thinplate_R2_node")
    P, Q = g["P"], g["Q"]
    reg = _reg(gpr, ctx, g)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"])
    fb, vb, gb = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    res = [None] * len(Q)

    def work(i):
        if i % 3 == 0:
            res[i] = reg.evaluate(m, Q[i:i + 1, 0], Q[i:i + 1, 1], Q[i:i + 1, 2], var=True, grad=True)
        elif i % 3 == 1:
            res[i] = reg.evaluate(m, Q[i:i + 1, 0], Q[i:i + 1, 1], Q[i:i + 1, 2], var=True) + (None,)
        else:
            res[i] = (reg.evaluate(m, Q[i:i + 1, 0], Q[i:i + 1, 1], Q[i:i + 1, 2]), None, None)

    for rep in range(3):
        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(Q))]
        [t.start() for t in threads]
        [t.join() for t in threads]
        for i, (f, v, gr) in enumerate(res):
            assert abs(f[0] - fb[i]) <= 1e-12 * np.abs(fb).max()
            if v is not None:
                assert abs(v[0] - vb[i]) <= 1e-10 * np.abs(vb).max()
            if gr is not None:
                assert np.abs(gr[0] - gb[i]).max() <= 1e-12 * np.abs(gb).max()
    assert relerr(fb, g["f"]) <= TOL_MEAN and np.abs(vb - g["v"]).max() <= TOL_VAR * np.abs(g["v"]).max()


def test_unchanged_caller_fanout_bench_against_the_reference_arm(gpr, tmp_path):
    """tests/cpp/fanout_bench.cpp — the node's loop verbatim (29 slabs x 841 std::threads x one evaluate(q = 1)) —
    through the drop-in headers, beside the SAME source compiled against the reference's own header
    (oracle/_ref/fanout_ref, prebuilt where /root/reference exists): identical kept set, mean / variance within the
    north-star tolerances at every lattice point, and the drop-in is not slower than the reference's CPU path."""
    sys_path = os.path.join(ROOT, "tools")
    import importlib.util
    spec = importlib.util.spec_from_file_location("fanout_bench", os.path.join(sys_path, "fanout_bench.py"))
    fb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fb)
    if not os.path.exists(fb.REF_EXE):
        pytest.skip("oracle/_ref/fanout_ref was not prebuilt")
    exe = fb.build_ours(str(tmp_path))
    for name, setting in (("mugD", "node"), ("mugD", "spd")):
        res, _ = fb.run_case(name, setting, exe, str(tmp_path))
        print(res)
        par = res["parity"]
        assert par["mean_rel_inf"] <= TOL_MEAN and par["var_rel_inf"] <= TOL_VAR and par["sign_mismatches"] == 0
        assert par["kept_equal"]
        # timing is reported by bench.py (fanout) and discussed in DESIGN.md: both arms are bound by the caller's own thread
        # spawn + join (~16-19 ms per slab of 841 threads); here only a coarse guard against a regression
        assert res["ours"]["total_s"] <= 2.0 * res["reference"]["total_s"], res


def test_batched_sample_on_chart_matches_the_atlas_loop(gpr, orc, ctx):
    """AtlasVariance::sampleOnChart (include/atlas/atlas_variance.hpp:147-219) for many charts in one call,
    against the reference's loop transcribed over single-query evaluations of the CPU ORACLE: same annulus points
    (pK = Tkl * pL, :166-195) from the same uniform variates, same mean / variance, same order by decreasing variance
    (vars_ids after the sort, :214-218).  Charts are built like createNode does (:68-105): gradient and variance at the
    centre, computeTangentBasis, radius = -0.3 v + 0.2 (:65, src/gp_node.cpp:959), ceil(200 R) samples."""
    g = load_golden("ref_mugD_thinplate_R2_node")                 # the node's own setting (indefinite K)
    P = g["P"]
    reg = _reg(gpr, ctx, g)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"])
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], g["y"], g["s2"], "thin_plate", g["p0"], 0.0, factor="ldlt")
    centres = P[:262:23]                                            # 12 surface points of the cloud
    fc, vc, gc = o.predict(centres[:, 0], centres[:, 1], centres[:, 2], var=True, grad=True)
    N, Tx, Ty = orc.tangent_basis(gc)
    R = -0.3 * vc + 0.2
    counts = np.ceil(200 * R).astype(np.int64)
    frames = np.hstack([centres, N, Tx, Ty, R[:, None]])
    rng = np.random.default_rng(3)
    total = int(counts.sum())
    r, th = rng.uniform(0.8, 1.0, total), rng.uniform(0.0, 2 * np.pi, total)
    pts, f, v, order = reg.sample_charts(m, frames, counts, r, th)
    off = np.concatenate([[0], np.cumsum(counts)])
    for c in range(len(centres)):
        sl = slice(off[c], off[c + 1])
        a = R[c] * np.sqrt(r[sl]) * np.cos(th[sl])
        b = R[c] * np.sqrt(r[sl]) * np.sin(th[sl])
        pk = Tx[c][None, :] * a[:, None] + Ty[c][None, :] * b[:, None] + N[c][None, :] * 0.0 + centres[c][None, :]
        assert np.abs(pts[sl] - pk).max() <= 1e-14
        fo = np.zeros(counts[c]); vo = np.zeros(counts[c])
        for i in range(counts[c]):                                  # one evaluate(f, v) per sample, like :201
            fi, vi, _ = o.predict(pk[i:i + 1, 0], pk[i:i + 1, 1], pk[i:i + 1, 2], var=True)
            fo[i], vo[i] = fi[0], vi[0]
        assert np.abs(f[sl] - fo).max() <= TOL_MEAN * max(np.abs(fo).max(), 1.0)
        assert np.abs(v[sl] - vo).max() <= TOL_VAR * np.abs(vo).max()
        ids = order[sl]
        assert sorted(ids.tolist()) == list(range(counts[c]))       # a permutation of the chart's samples
        assert (np.diff(v[sl][ids]) <= 0).all()                     # decreasing in our own variances, exactly
        ref_ids = sorted(range(counts[c]), key=lambda i: (-vo[i], i))
        tol = 2 * TOL_VAR * np.abs(vo).max()
        assert (np.abs(vo[ids] - vo[ref_ids]) <= tol).all()         # same order as the reference loop up to ties within tolerance
        assert ids[0] == ref_ids[0] or abs(vo[ids[0]] - vo[ref_ids[0]]) <= tol       # the sample getNextState picks (:121-122)
    # library-drawn variates: deterministic per seed, inside the annulus, in the tangent plane
    p1, f1, v1, o1 = reg.sample_charts(m, frames, counts, seed=7)
    p2, f2, v2, o2 = reg.sample_charts(m, frames, counts, seed=7)
    p3 = reg.sample_charts(m, frames, counts, seed=8)[0]
    assert np.array_equal(p1, p2) and np.array_equal(v1, v2) and np.array_equal(o1, o2) and not np.array_equal(p1, p3)
    for c in range(len(centres)):
        d = p1[off[c]:off[c + 1]] - centres[c]
        rad = np.linalg.norm(d, axis=1)
        assert (rad >= np.sqrt(0.8) * R[c] * (1 - 1e-12)).all() and (rad <= R[c] * (1 + 1e-12)).all()
        assert np.abs(d @ N[c]).max() <= 1e-12


def _oracle_marching(o, lo, hi, step, leaf, pas, tol):
    """src/gp_node.cpp:1103-1292 transcribed (start search :1124-1150, marchingCubes :1195-1292) with the CPU oracle answering
    the evaluate calls; float32 coordinate arithmetic like pcl::PointXYZ; cubes identified by integer offsets (the reference's
    octree voxel test), cube centres by repeated float additions of the leaf size."""
    f32 = np.float32
    leaf, pas = f32(leaf), f32(pas)
    axis, a = [], lo
    while a <= hi:
        axis.append(a)
        a += step
    G = np.array([(x, y, z) for x in axis for y in axis for z in axis])
    fl = o.predict(G[:, 0], G[:, 1], G[:, 2], threads=os.cpu_count() or 1)[0]
    hit = np.flatnonzero(np.abs(fl) <= tol)
    assert len(hit)
    start = G[hit[0]].astype(f32)
    steps = int(np.round(leaf / pas))

    def centre(c):
        out = []
        for ax in range(3):
            v = start[ax]
            for _ in range(abs(c[ax])):
                v = f32(v + leaf) if c[ax] > 0 else f32(v - leaf)
            out.append(v)
        return out

    visited, frontier, kept, cubes = {(0, 0, 0)}, [(0, 0, 0)], {}, 0
    while frontier:
        nxt = []
        for c in sorted(frontier):
            cubes += 1
            ctr = centre(c)
            pts, ijk = [], []
            for i in range(steps + 1):
                for j in range(steps + 1):
                    for k in range(steps + 1):
                        pts.append([float(f32(f32(ctr[0] - leaf / f32(2)) + f32(f32(i) * pas))),
                                    float(f32(f32(ctr[1] - leaf / f32(2)) + f32(f32(j) * pas))),
                                    float(f32(f32(ctr[2] - leaf / f32(2)) + f32(f32(k) * pas)))])
                        ijk.append((i, j, k))
            pts = np.array(pts)
            ff, vv, _ = o.predict(pts[:, 0], pts[:, 1], pts[:, 2], var=True, threads=os.cpu_count() or 1)
            where = [False] * 6
            for p, (i, j, k), fi, vi in zip(pts, ijk, ff, vv):
                if abs(fi) <= tol:
                    for d, cond in enumerate((i == 0, i == steps, j == 0, j == steps, k == 0, k == steps)):
                        where[d] = where[d] or cond
                    kept.setdefault((c[0] * steps + i, c[1] * steps + j, c[2] * steps + k), (p, fi, vi))
            for d in range(6):
                if where[d]:
                    nb = list(c)
                    nb[d // 2] += 1 if d % 2 else -1
                    nb = tuple(nb)
                    if nb not in visited:
                        visited.add(nb)
                        nxt.append(nb)
        frontier = nxt
    return kept, cubes


def test_batched_marching_sampler_matches_the_node_loop(gpr, orc, ctx):
    """SURVEY §8(f).2, second half: the node's marchingSampling / marchingCubes (src/gp_node.cpp:1103-1292) as a wave-by-wave
    batched flood fill, against the reference loop transcribed over the CPU oracle: same starting point, same set of
    visited cubes, same kept samples (coordinates bit for bit, f and v within tolerance)."""
    g = load_golden("ref_mugD_lattice_spd")
    P, y, s2, R = g["P"], g["y"], g["s2"], float(g["p0"])
    reg = gpr.GPRegressor("thin_plate", R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", R, 0.0, factor="llt")
    kw = dict(lo=-1.1, hi=1.1, step=0.1, leaf=0.12, tol=0.01)          # a coarser leaf than the node's 0.06 keeps the CPU loop short
    pts, f, v, cubes = reg.sample_marching(m, leaf_pass=0.04, **kw)
    kept, ref_cubes = _oracle_marching(o, kw["lo"], kw["hi"], kw["step"], kw["leaf"], 0.04, kw["tol"])
    fo = np.array([k[1] for k in kept.values()])
    edge = np.abs(np.abs(fo) - 0.01) < 1e-9
    assert len(pts) > 200 and v.min() > 0
    if not edge.any():
        assert cubes == ref_cubes and len(pts) == len(kept)
        ref_pts = np.array([k[0] for k in kept.values()])
        order_g = np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0]))
        order_r = np.lexsort((ref_pts[:, 2], ref_pts[:, 1], ref_pts[:, 0]))
        assert np.array_equal(pts[order_g], ref_pts[order_r])
        vo = np.array([k[2] for k in kept.values()])
        assert np.abs(f[order_g] - fo[order_r]).max() <= 1e-9
        assert np.abs(v[order_g] - vo[order_r]).max() <= TOL_VAR * np.abs(vo).max()
    # the node's own parameters (leaf 0.06, pass 0.02, :258): runs, stays on the surface, bounded output
    pts2, f2, v2, cubes2 = reg.sample_marching(m)
    assert len(pts2) > len(pts) and np.abs(f2).max() <= 0.01 and cubes2 > cubes
    with pytest.raises(gpr.GPRegressionException, match="No starting point"):
        reg.sample_marching(m, tol=1e-12)


@pytest.mark.parametrize("case,batches", [("ref_mugD_thinplate_R2_node", (3, 5, 30)), ("ref_jug_thinplate_R2_node", (33, 30, 20))])
def test_incremental_update_of_an_indefinite_tail_model(gpr, orc, ctx, case, batches):
    """update<>() on the node's own configuration (ThinPlate(2.0), 15 external points: indefinite K).  The reference's
    update refits with the pivoted LDLT (gp_regressor.hpp:442-459); here the touch points are appended to the positive
    definite leading block and the trailing block is eliminated again (no refit: append_ms > 0).  Checked against the
    pivoted-LDLT oracle (which refits, like the reference) after every batch, and against a fresh fit; crosses a tile
    boundary and a capacity growth."""
    g = load_golden(case)
    P, y, s2, Q = g["P"], g["y"], g["s2"], g["Q"]
    reg = _reg(gpr, ctx, g)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    assert m.n_tail == 15
    o = orc.Oracle(P[:, 0], P[:, 1], P[:, 2], y, s2, "thin_plate", g["p0"], 0.0, factor="ldlt")
    rng = np.random.default_rng(len(P))
    allP, ally, alls = [P], [y], [s2]
    for k in batches:
        idx = rng.choice(len(P) - 15, size=k, replace=False)
        Pn = P[idx] * (1.0 + 0.02 * rng.standard_normal((k, 1))) + 0.01 * rng.standard_normal((k, 3))   # touches near the surface
        yn, sn = np.zeros(k), np.full(k, 0.05)
        reg.update(m, Pn[:, 0], Pn[:, 1], Pn[:, 2], yn, sn)
        assert ctx.timings()["append_ms"] > 0.0 and m.n_tail == 15
        o.update(Pn[:, 0], Pn[:, 1], Pn[:, 2], yn, sn)
        allP.append(Pn); ally.append(yn); alls.append(sn)
        assert relerr(m.alpha, o.alpha) <= TOL_ALPHA
        f, v, gr = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
        fo, vo, go = o.predict(Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True, threads=4)
        assert relerr(f, fo) <= TOL_MEAN and relerr(gr, go) <= TOL_MEAN and _signs_agree(f, fo)
        assert np.abs(v - vo).max() <= TOL_VAR * np.abs(vo).max()
        f1, v1 = reg.evaluate(m, Q[:1, 0], Q[:1, 1], Q[:1, 2], var=True)                        # fused single-query kernel
        assert abs(f1[0] - fo[0]) <= TOL_MEAN * np.abs(fo).max() and abs(v1[0] - vo[0]) <= TOL_VAR * np.abs(vo).max()
    A, ya, sa = np.vstack(allP), np.concatenate(ally), np.concatenate(alls)
    fresh = reg.create(A[:, 0], A[:, 1], A[:, 2], ya, sa)
    assert fresh.n == m.n and fresh.n_tail == 15 and relerr(m.alpha, fresh.alpha) <= TOL_ALPHA
    # a point that cannot join the leading block (far outside, farther than R from the cloud): falls back to the refit,
    # which moves it into the trailing block
    far = np.array([[0.0, 0.0, 2.6]])
    reg.update(m, far[:, 0], far[:, 1], far[:, 2], np.ones(1), np.full(1, 0.1))
    o.update(far[:, 0], far[:, 1], far[:, 2], np.ones(1), np.full(1, 0.1))
    assert m.n_tail >= 16 and relerr(m.alpha, o.alpha) <= TOL_ALPHA


def test_model_save_and_load_round_trip(gpr, ctx, tmp_path):
    """Export / import: with the stored factor the loaded model answers without refactorising
    (same alpha and factor bits); without it (and for an indefinite-tail model) it is refitted from the stored
    training set."""
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(700, seed=13)
    reg = gpr.GPRegressor("gaussian", 1.3, 0.9, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2, with_normals=True)
    Q = W.grid_slab(10, 3, 5)
    ref = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
    one = reg.evaluate(m, Q[:1, 0], Q[:1, 1], Q[:1, 2], var=True)
    for with_factor in (True, False):
        path = tmp_path / ("model_%d.bin" % with_factor)
        reg.save(m, path, with_factor)
        assert os.path.getsize(path) > (700 * 701 // 2 * 8 if with_factor else 0)
        other = gpr.GPRegressor("thin_plate", 1.0, ctx=ctx)             # the kernel comes from the file
        m2 = other.load(path, with_normals=True)
        assert m2.n == m.n and m2.R == m.R and np.array_equal(m2.alpha, m.alpha)
        assert np.array_equal(m2.get()["normals"], m.get()["normals"])
        got = other.evaluate(m2, Q[:, 0], Q[:, 1], Q[:, 2], var=True, grad=True)
        # mean and gradient depend on alpha only: identical bits.  The variance goes through L^-1, rebuilt from the
        # loaded factor with re-derived diagonal-block inverses (1/L_ii instead of the fit's rsqrt): last bits may differ.
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2])
        assert np.abs(got[1] - ref[1]).max() <= 1e-11 * np.abs(ref[1]).max()
        got1 = other.evaluate(m2, Q[:1, 0], Q[:1, 1], Q[:1, 2], var=True)
        assert np.array_equal(got1[0], one[0]) and abs(got1[1][0] - one[1][0]) <= 1e-11 * np.abs(ref[1]).max()
        if with_factor:
            assert np.array_equal(m2.factor(), m.factor())
            other.update(m2, P[:5, 0] * 0.5, P[:5, 1] * 0.5, P[:5, 2] * 0.5, y[:5], s2[:5])     # a loaded model can be appended to
            assert m2.n == m.n + 5
    g = load_golden("ref_mugD_thinplate_R2_node")                         # indefinite tail: refit on load
    Pn = g["P"]
    regn = _reg(gpr, ctx, g)
    mn = regn.create(Pn[:, 0], Pn[:, 1], Pn[:, 2], g["y"], g["s2"])
    regn.save(mn, tmp_path / "node.bin")
    mn2 = regn.load(tmp_path / "node.bin")
    assert mn2.n_tail == 15 and np.array_equal(mn2.alpha, mn.alpha)
    with pytest.raises(gpr.GPRegressionException):
        regn.load(tmp_path / "missing.bin")


def test_sampler_sharded_over_two_devices_is_identical(gpr):
    """One context over two GPUs: the sampler splits the lattice into contiguous ranges, one per device; the points, their
    order and every bit of f must equal the single-device result (each lattice point is an independent query)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(1200, seed=3)
    outs = []
    for devs in ([0], [0, 1]):
        c = gpr.Context(devices=devs)
        reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=c)
        m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
        outs.append(reg.sample_isosurface(m, lo=-1.2, hi=1.2, step=2.4 / 139, tol=0.002))     # 140^3 = 2.7 M points
        # the batched projection is sharded the same way (one CTA per point, contiguous ranges per device)
        rng = np.random.default_rng(4)
        d = rng.standard_normal((300, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        start = d * rng.uniform(0.9, 1.2, size=(300, 1))
        _, g0 = reg.evaluate(m, start[:, 0], start[:, 1], start[:, 2], grad=True)
        po, ps = reg.project(m, start, g0, f_tol=1e-3, improve_tol=1e-9, max_iter=60, step_mul=0.2)
        outs[-1] = outs[-1] + (po, ps)
        del m
        c.close()
    assert len(outs[0][0]) > 100
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    # a context whose primary device is GPU 1, after GPU 0 has run everything: per-device kernel attributes
    # (dynamic shared memory limits are set per device, not per process)
    c1 = gpr.Context(devices=[1])
    c0 = gpr.Context(devices=[0])
    Q = W.grid_slab(16, 0, 16)
    res = []
    for c in (c0, c1):
        reg = gpr.GPRegressor("thin_plate", W.SYNTH_R, ctx=c)
        m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
        f, v = reg.evaluate(m, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
        reg.update(m, P[:8, 0] * 1.01, P[:8, 1] * 1.01, P[:8, 2] * 1.01, y[:8], s2[:8])
        res.append((m.alpha.copy(), f, v))
        del m
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_tail_block_is_replicated_across_processes(gpr):
    """Two ranks under torchrun (needs 2 GPUs): the indefinite-tail model of the node's configuration is broadcast with
    NCCL and both ranks reproduce the reference fixture on their query shards."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([os.sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", os.path.join(ROOT, "tools", "tail_replica_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "TAIL_REPLICA_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_factor_is_published_to_replicas_during_the_fit(gpr):
    """Two ranks under torchrun (needs 2 GPUs): the replication fused into the Cholesky kernel (peer stores through CUDA IPC
    mappings, distributed.fit_and_publish) — sharded results bit-identical to rank 0's own, identical to the legacy
    broadcast path, and the indefinite configuration falls back to the broadcast."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([os.sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29543", os.path.join(ROOT, "tools", "publish_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "PUBLISH_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]

"""ctypes bindings and numpy/scipy flavour of the CPU oracle (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libgpr_ref.so")
_dp = C.POINTER(C.c_double)

KINDS = {"thin_plate": 0, "gaussian": 1, "laplace": 2}


def build(force=False):
    """Compile liboracle.so (and _ref/libgpr_ref.so when /root/reference is present)."""
    src = os.path.join(_HERE, "gpr_oracle.cpp")
    stale = (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < os.path.getmtime(src)
    if force or stale or (os.path.isdir("/root/reference/include") and not os.path.exists(_REF)):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


def have_reference():
    return os.path.exists(_REF)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_fit.restype = C.c_void_p
        L.orc_fit.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double,
                              C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_n.argtypes = [C.c_void_p]
        L.orc_info.argtypes = [C.c_void_p]
        L.orc_R.argtypes = [C.c_void_p]
        L.orc_R.restype = C.c_double
        L.orc_get.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.orc_predict.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp, C.c_int]
        L.orc_update.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_int]
        L.orc_tangent_basis.argtypes = [_dp, C.c_int, _dp, _dp, _dp]
        L.orc_kernel.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_kernel.restype = C.c_double
        L.orc_residual.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_int, _dp, C.c_int]
        _lib = L
    return _lib


def kernel_value(kind, p0, p1, d, diff=False):
    return _load().orc_kernel(KINDS[kind], float(p0), float(p1), float(d), int(diff))


def residual(P, sigma2, kind, p0, p1, B, W, threads=None):
    """R = B - K W in extended precision (K never stored; oracle/gpr_oracle.cpp: orc_residual).  B, W: (n, m)."""
    P = np.asarray(P, dtype=np.float64)
    n = P.shape[0]
    Bf = np.asfortranarray(np.asarray(B, dtype=np.float64).reshape(n, -1))
    Wf = np.asfortranarray(np.asarray(W, dtype=np.float64).reshape(n, -1))
    R = np.zeros_like(Bf, order="F")
    x, y, z, s2 = (_f64(P[:, 0]), _f64(P[:, 1]), _f64(P[:, 2]), _f64(sigma2))
    _load().orc_residual(_p(x), _p(y), _p(z), _p(s2), n, KINDS[kind], float(p0), float(p1), _p(Bf), _p(Wf), Bf.shape[1],
                         _p(R), int(threads or os.cpu_count() or 1))
    return np.ascontiguousarray(R)


def certify(P, y, sigma2, kind, p0, p1, Q, alpha, W, threads=None):
    """A-posteriori certificate of mean and variance at queries Q (q,3) from ANY approximate solves
    alpha ~ K^-1 y and W[:, j] ~ K^-1 k*_j (n, q):  with r = k* - K w and r_a = y - K alpha (extended precision),
        v = k(0) - k*^T w - w^T r - r^T K^-1 r,      f = k*^T alpha + w^T r_a + (second order),
    identities that hold exactly for symmetric K (gp_regressor.hpp:318-319, :305 restated).  Returns a dict with
    f, v, the neglected second-order bounds (|r|^2 / min sigma2: K = PSD kernel + diag(sigma2)) and the residual norms."""
    P = np.asarray(P, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    Ks = _kern(kind, p0, p1, _pdist(P, Q))                         # (n, q): columns k*_j
    B = np.concatenate([Ks, np.asarray(y, dtype=np.float64)[:, None]], axis=1)
    Wa = np.concatenate([np.asarray(W, dtype=np.float64), np.asarray(alpha, dtype=np.float64)[:, None]], axis=1)
    R = residual(P, sigma2, kind, p0, p1, B, Wa, threads)
    ld = np.longdouble
    k0 = float(_kern(kind, p0, p1, np.zeros(1))[0])
    Wl, Rl, Kl = Wa[:, :-1].astype(ld), R[:, :-1].astype(ld), Ks.astype(ld)
    v = ld(k0) - (Kl * Wl).sum(axis=0) - (Wl * Rl).sum(axis=0)
    f = (Kl * Wa[:, -1:].astype(ld)).sum(axis=0) + (Wl * R[:, -1:].astype(ld)).sum(axis=0)
    lam = float(np.min(sigma2)) if sigma2 is not None else None
    r2 = (Rl * Rl).sum(axis=0)
    return {"f": f.astype(np.float64), "v": v.astype(np.float64), "Ks": Ks, "R": R,
            "v_second_order": None if lam is None else float(r2.max()) / lam,
            "f_second_order": None if lam is None else float(np.sqrt(r2.max() * (R[:, -1].astype(ld) ** 2).sum())) / lam,
            "resid_inf": float(np.abs(R[:, :-1]).max()), "resid_alpha_inf": float(np.abs(R[:, -1]).max())}


def tangent_basis(grad):
    """grad: (q,3) -> N, Tx, Ty each (q,3).  gp_regressor.hpp:29-44, :204-211."""
    g = np.asfortranarray(np.asarray(grad, dtype=np.float64).reshape(-1, 3))
    q = g.shape[0]
    out = [np.zeros((q, 3), order="F") for _ in range(3)]
    _load().orc_tangent_basis(_p(g), q, *[_p(o) for o in out])
    return tuple(np.ascontiguousarray(o) for o in out)


class Oracle:
    """Our restatement.  factor: 'ldlt' (reference) | 'llt';  dist: 'diff' | 'expansion';
    precision: 'double' | 'longdouble'."""

    def __init__(self, x, y, z, label, sigma2, kind="thin_plate", p0=1.0, p1=1.0, factor="ldlt",
                 dist="diff", with_normals=False, precision="double"):
        L = _load()
        x, y, z, label, sigma2 = map(_f64, (x, y, z, label, sigma2))
        self._h = L.orc_fit(_p(x), _p(y), _p(z), _p(label), _p(sigma2), len(x), KINDS[kind], float(p0),
                            float(p1), 0 if factor == "ldlt" else 1, 0 if dist == "diff" else 1,
                            int(with_normals), 0 if precision == "double" else 1)
        self.with_normals = with_normals

    def __del__(self):
        if getattr(self, "_h", None):
            _load().orc_free(self._h)
            self._h = None

    @property
    def n(self):
        return _load().orc_n(self._h)

    @property
    def info(self):
        """0, or 1+index of the first non-positive pivot (LLT only)."""
        return _load().orc_info(self._h)

    @property
    def R(self):
        return _load().orc_R(self._h)

    def get(self, K=False, factor=False):
        n = self.n
        alpha = np.zeros(n)
        N = np.zeros((n, 3), order="F") if self.with_normals else None
        Km = np.zeros((n, n), order="F") if K else None
        Fm = np.zeros((n, n), order="F") if factor else None
        _load().orc_get(self._h, _p(alpha), _p(N), _p(Km), _p(Fm))
        return {"alpha": alpha, "normals": None if N is None else np.ascontiguousarray(N), "K": Km, "factor": Fm}

    @property
    def alpha(self):
        return self.get()["alpha"]

    def predict(self, qx, qy, qz, var=False, grad=False, threads=1):
        qx, qy, qz = map(_f64, (qx, qy, qz))
        q = len(qx)
        f = np.zeros(q)
        v = np.zeros(q) if var else None
        g = np.zeros((q, 3), order="F") if grad else None
        _load().orc_predict(self._h, _p(qx), _p(qy), _p(qz), q, _p(f), _p(v), _p(g), int(threads))
        return f, v, (None if g is None else np.ascontiguousarray(g))

    def update(self, x, y, z, label, sigma2):
        x, y, z, label, sigma2 = map(_f64, (x, y, z, label, sigma2))
        _load().orc_update(self._h, _p(x), _p(y), _p(z), _p(label), _p(sigma2), len(x))


class Reference:
    """The reference's own gp_regressor.hpp (compiled against oracle/eigen_shim)."""

    def __init__(self, kind="thin_plate", p0=1.0, p1=1.0):
        if not have_reference():
            raise RuntimeError("oracle/_ref/libgpr_ref.so missing: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(_REF)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_double, C.c_double]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_error.argtypes = [C.c_void_p]
        L.ref_error.restype = C.c_char_p
        L.ref_fit.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int]
        L.ref_n.argtypes = [C.c_void_p]
        L.ref_R.argtypes = [C.c_void_p]
        L.ref_R.restype = C.c_double
        L.ref_get.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.ref_evaluate.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.ref_update.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_int]
        L.ref_adopt.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, C.c_double]
        L.ref_evaluate_mt.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.ref_error_message.argtypes = [C.c_int]
        L.ref_error_message.restype = C.c_char_p
        self._L = L
        self._h = L.ref_create(KINDS[kind], float(p0), float(p1))
        self.with_normals = False

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_free(self._h)
            self._h = None

    def _check(self, rc):
        if rc:
            raise RuntimeError(self._L.ref_error(self._h).decode())

    def fit(self, x, y, z, label, sigma2, with_normals=False):
        x, y, z, label, sigma2 = map(_f64, (x, y, z, label, sigma2))
        self.with_normals = with_normals
        self._check(self._L.ref_fit(self._h, _p(x), _p(y), _p(z), _p(label), _p(sigma2), len(x), int(with_normals)))
        return self

    @property
    def n(self):
        return self._L.ref_n(self._h)

    @property
    def R(self):
        return self._L.ref_R(self._h)

    def get(self, K=False):
        n = self.n
        alpha = np.zeros(n)
        N = np.zeros((n, 3), order="F") if self.with_normals else None
        Km = np.zeros((n, n), order="F") if K else None
        self._L.ref_get(self._h, _p(alpha), _p(N), _p(Km))
        return {"alpha": alpha, "normals": None if N is None else np.ascontiguousarray(N), "K": Km}

    def evaluate(self, qx, qy, qz, mode=2):
        """mode 1: f; 2: f,v; 3: f,v,grad; 4: f,v,grad,Tx,Ty (the four reference overloads)."""
        qx, qy, qz = map(_f64, (qx, qy, qz))
        q = len(qx)
        f, v = np.zeros(q), np.zeros(q)
        N, Tx, Ty = (np.zeros((q, 3), order="F") for _ in range(3))
        self._check(self._L.ref_evaluate(self._h, _p(qx), _p(qy), _p(qz), q, mode, _p(f), _p(v), _p(N), _p(Tx), _p(Ty)))
        out = [f]
        if mode >= 2:
            out.append(v)
        if mode >= 3:
            out.append(np.ascontiguousarray(N))
        if mode >= 4:
            out += [np.ascontiguousarray(Tx), np.ascontiguousarray(Ty)]
        return tuple(out)

    def adopt(self, x, y, z, label, sigma2, alpha, Lc, R):
        """Timing support (bench.py --impl reference): install a model whose factor was computed elsewhere
        (Lc: n x n lower Cholesky factor) so that the reference's own evaluate() can run at sizes where its
        unblocked LDLT::compute would take tens of minutes."""
        x, y, z, label, sigma2, alpha = map(_f64, (x, y, z, label, sigma2, alpha))
        Lc = np.asfortranarray(Lc, dtype=np.float64)
        self._check(self._L.ref_adopt(self._h, _p(x), _p(y), _p(z), _p(label), _p(sigma2), len(x), _p(alpha), _p(Lc), float(R)))
        return self

    def evaluate_mt(self, qx, qy, qz, var=True, threads=1, per_call=1):
        """The reference's evaluate() from `threads` concurrent threads, `per_call` queries per call."""
        qx, qy, qz = map(_f64, (qx, qy, qz))
        q = len(qx)
        f, v = np.zeros(q), np.zeros(q)
        self._check(self._L.ref_evaluate_mt(self._h, _p(qx), _p(qy), _p(qz), q, 2 if var else 1, int(threads), int(per_call), _p(f), _p(v)))
        return f, (v if var else None)

    def update(self, x, y, z, label, sigma2):
        x, y, z, label, sigma2 = map(_f64, (x, y, z, label, sigma2))
        self._check(self._L.ref_update(self._h, _p(x), _p(y), _p(z), _p(label), _p(sigma2), len(x)))

    def error_message(self, which):
        return self._L.ref_error_message(which).decode()


# ----------------------------------------------------------------------------------------------
# numpy / scipy flavour ("best-effort CPU"): same mathematics, LAPACK on all
# host cores.  Follows gp_regressor.hpp:110-163 (fit) and :282-324 (mean + variance).
# ----------------------------------------------------------------------------------------------
def _kern(kind, p0, p1, d):
    if kind == "thin_plate":
        return 2 * d ** 3 - 3 * p0 * d ** 2 + p0 ** 3          # kernels/thin_plate.hpp:14
    if kind == "gaussian":
        return (p0 * p0) * np.exp(-d / (p1 * p1))              # kernels/gaussian.hpp:17-18
    return 2 * p0 * np.exp(-d / p1)                            # kernels/laplace.hpp:39-40


def _pdist(A, B):
    d2 = np.zeros((A.shape[0], B.shape[0]))
    for c in range(3):
        diff = A[:, c:c + 1] - B[None, :, c]
        d2 += diff * diff
    return np.sqrt(d2)


def blas_fit(P, y, sigma2, kind="thin_plate", p0=1.0, p1=1.0):
    """Returns dict(L, alpha, P): K = k(D) + diag(sigma2); L = chol(K) (dpotrf); alpha = K^-1 y."""
    from scipy.linalg import cho_solve, cholesky
    P = np.asarray(P, dtype=np.float64)
    K = _kern(kind, p0, p1, _pdist(P, P))
    if sigma2 is not None:
        K[np.diag_indices_from(K)] += sigma2
    L = cholesky(K, lower=True, overwrite_a=True, check_finite=False)
    alpha = cho_solve((L, True), np.asarray(y, dtype=np.float64), check_finite=False)
    return {"L": L, "alpha": alpha, "P": P, "kind": kind, "p0": p0, "p1": p1}


def blas_predict(model, Q, var=True):
    """mean (and variance) for queries Q (q,3): f = K* alpha; v = k(0) - |L^-1 K*^T|^2 (dtrsm)."""
    from scipy.linalg import solve_triangular
    Ks = _kern(model["kind"], model["p0"], model["p1"], _pdist(np.asarray(Q, dtype=np.float64), model["P"]))
    f = Ks @ model["alpha"]
    if not var:
        return f, None
    V = solve_triangular(model["L"], Ks.T, lower=True, check_finite=False, overwrite_b=True)
    k0 = float(_kern(model["kind"], model["p0"], model["p1"], np.zeros(1))[0])
    return f, k0 - np.einsum("ij,ij->j", V, V)

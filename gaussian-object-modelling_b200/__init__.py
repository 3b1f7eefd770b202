"""gaussian-object-modelling_b200 — B200-native GP-regression core (hot path of
pacman-project/gaussian-object-modelling, include/gp_regression).

The product is ``libgpr_b200.so`` (hand-written sm_100a CUDA behind the C-ABI in include/gpr_c_api.h)
plus the C++ drop-in headers in include/gp_regression/.  This Python module is only the ctypes binding
that tests/ and bench.py use to call the C-ABI; it mirrors the reference's regressor interface
(create / evaluate / update over Data-like SoA arrays, gp_regressor.hpp:110,:194-357,:367) and its error
behaviour.  It never computes anything itself and never imports the CPU oracle: if the CUDA library is
missing or no GPU is usable, calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpr_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

KINDS = {"thin_plate": 0, "gaussian": 1, "laplace": 2}
GPR_OK, GPR_ERR_INVALID, GPR_ERR_NOT_SPD, GPR_ERR_CUDA, GPR_ERR_OOM = range(5)

_dp = C.POINTER(C.c_double)


class GPRegressionException(Exception):
    """Counterpart of gp_regression::GPRegressionException (gp_regression_exception.h:9-17)."""

    def __init__(self, message, code=GPR_ERR_INVALID, pivot=0):
        super().__init__(message)
        self.code = code
        self.pivot = pivot


class KernelT(C.Structure):
    _fields_ = [("kind", C.c_int), ("p0", C.c_double), ("p1", C.c_double)]


class Timings(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "cov_ms", "chol_ms", "solve_ms", "normals_ms", "fit_total_ms", "linv_ms",
        "predict_mean_ms", "predict_var_ms", "predict_total_ms", "h2d_ms", "d2h_ms", "append_ms", "ozaki_ms", "ozaki_slices", "ozaki_issued_fraction",
        "fit_int8_slices")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ModelState(C.Structure):
    _fields_ = [("n", C.c_size_t), ("padded_n", C.c_size_t), ("ld", C.c_size_t), ("kernel", KernelT), ("R", C.c_double),
                ("xyz", C.c_void_p), ("alpha", C.c_void_p), ("linv", C.c_void_p),
                ("n_tail", C.c_size_t), ("tail_pad", C.c_size_t), ("tail_z", C.c_void_p), ("tail_sinv", C.c_void_p),
                ("lfac", C.c_void_p), ("dinv", C.c_void_p)]


def build(force=False):
    """Compile libgpr_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    cmd = ["make", "-C", csrc, "-j8"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it was not built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libgpr_b200.so is missing (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, sz, ci, cd = C.c_void_p, C.c_size_t, C.c_int, C.c_double
        L.gpr_ctx_create.argtypes = [C.POINTER(ci), ci, C.POINTER(vp)]
        L.gpr_ctx_destroy.argtypes = [vp]
        L.gpr_ctx_num_devices.argtypes = [vp]
        L.gpr_last_error.restype = C.c_char_p
        L.gpr_last_pivot.restype = C.c_longlong
        L.gpr_last_timings.argtypes = [vp, C.POINTER(Timings)]
        L.gpr_fit.argtypes = [vp, _dp, _dp, _dp, _dp, _dp, sz, KernelT, ci, C.POINTER(vp)]
        L.gpr_model_destroy.argtypes = [vp]
        L.gpr_model_size.argtypes = [vp]
        L.gpr_model_size.restype = sz
        L.gpr_model_tail_size.argtypes = [vp]
        L.gpr_model_tail_size.restype = sz
        L.gpr_model_get.argtypes = [vp, _dp, _dp, _dp]
        L.gpr_model_get_factor.argtypes = [vp, _dp]
        L.gpr_predict.argtypes = [vp, vp, _dp, _dp, _dp, sz, _dp, _dp, _dp, _dp, _dp]
        L.gpr_predict_device.argtypes = [vp, vp, vp, vp, vp, sz, vp, vp, vp]
        L.gpr_sample_isosurface.argtypes = [vp, vp, cd, cd, cd, cd, sz, _dp, _dp, _dp, _dp, _dp, C.POINTER(sz)]
        L.gpr_project.argtypes = [vp, vp, _dp, _dp, _dp, _dp, _dp, _dp, sz, cd, cd, C.c_uint, cd, _dp, _dp, _dp, C.POINTER(ci)]
        L.gpr_sample_marching.argtypes = [vp, vp, cd, cd, cd, C.c_float, C.c_float, cd, sz, _dp, _dp, _dp, _dp, _dp, C.POINTER(sz), C.POINTER(sz)]
        L.gpr_sample_chart.argtypes = [vp, vp, _dp, C.POINTER(sz), sz, _dp, _dp, C.c_ulonglong, _dp, _dp, _dp, _dp, _dp, C.POINTER(sz)]
        L.gpr_model_save.argtypes = [vp, vp, C.c_char_p, ci]
        L.gpr_model_load.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        L.gpr_model_prepare_variance.argtypes = [vp, vp]
        L.gpr_model_solve.argtypes = [vp, vp, _dp, sz, _dp]
        L.gpr_model_ipc_export.argtypes = [vp, vp, vp]
        L.gpr_ctx_set_fit_peers.argtypes = [vp, vp, ci, sz]
        L.gpr_ctx_clear_fit_peers.argtypes = [vp]
        L.gpr_ctx_last_fit_published.argtypes = [vp]
        L.gpr_pcd_read_xyz.argtypes = [C.c_char_p, C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp), C.POINTER(sz)]
        L.gpr_free.argtypes = [vp]
        L.gpr_free.restype = None
        L.gpr_append.argtypes = [vp, vp, _dp, _dp, _dp, _dp, _dp, sz]
        L.gpr_model_reserve.argtypes = [vp, vp, sz]
        L.gpr_model_state_get.argtypes = [vp, vp, ci, C.POINTER(ModelState)]
        L.gpr_model_create_replica.argtypes = [vp, sz, KernelT, cd, ci, C.POINTER(vp)]
        L.gpr_model_create_replica_tail.argtypes = [vp, sz, sz, KernelT, cd, ci, C.POINTER(vp)]
        L.gpr_selftest_gemm.argtypes = [_dp, _dp, ci, _dp, ci, ci, ci]
        L.gpr_selftest_leaf.argtypes = [_dp, _dp, C.POINTER(ci)]
        L.gpr_selftest_factor.argtypes = [_dp, ci, _dp, ci, C.POINTER(C.c_longlong)]
        L.gpr_selftest_peak.argtypes = [ci, ci, _dp]
        L.gpr_selftest_i8gemm.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
        L.gpr_selftest_factor_trace.argtypes = [ci, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        _lib = L
    return _lib


# Every symbol include/gpr_c_api.h declares (checked by the CPU test-suite without a GPU).
C_ABI_SYMBOLS = [
    "gpr_ctx_create", "gpr_ctx_destroy", "gpr_ctx_num_devices", "gpr_last_error", "gpr_last_pivot",
    "gpr_last_timings", "gpr_fit", "gpr_model_destroy", "gpr_model_size", "gpr_model_tail_size", "gpr_model_get",
    "gpr_model_get_factor", "gpr_predict", "gpr_predict_device", "gpr_model_prepare_variance", "gpr_append",
    "gpr_model_reserve", "gpr_sample_isosurface", "gpr_project", "gpr_model_save", "gpr_model_load",
    "gpr_model_state_get", "gpr_model_create_replica", "gpr_model_create_replica_tail", "gpr_model_solve",
    "gpr_pcd_read_xyz", "gpr_free", "gpr_sample_chart", "gpr_sample_marching",
    "gpr_model_ipc_export", "gpr_ctx_set_fit_peers", "gpr_ctx_clear_fit_peers", "gpr_ctx_last_fit_published",
]
# Engine self-tests / pipe probes (csrc/gpr_selftest.h): exported for tests/ and bench.py, not part of the boundary.
SELFTEST_SYMBOLS = ["gpr_selftest_gemm", "gpr_selftest_leaf", "gpr_selftest_factor", "gpr_selftest_peak",
                    "gpr_selftest_factor_trace", "gpr_selftest_i8gemm"]


def _check(rc):
    if rc != GPR_OK:
        L = lib()
        raise GPRegressionException(L.gpr_last_error().decode(), rc, L.gpr_last_pivot() if rc == GPR_ERR_NOT_SPD else 0)


def _arr(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _same_length(*arrays):
    """The C side reads n entries of every array: refuse ragged input here, with the message of the C++ shim."""
    sizes = {a.size for a in arrays if a is not None}
    if len(sizes) > 1 or any(a is not None and a.ndim != 1 for a in arrays):
        raise GPRegressionException("Inconsistent input data sizes")


def pcd_read_xyz(path):
    """The C-ABI PCD reader (gpr_pcd_read_xyz): (n, 3) float64.  Host only, needs no GPU."""
    L = lib()
    px, py, pz, n = _dp(), _dp(), _dp(), C.c_size_t()
    _check(L.gpr_pcd_read_xyz(str(path).encode(), C.byref(px), C.byref(py), C.byref(pz), C.byref(n)))
    try:
        out = np.stack([np.ctypeslib.as_array(p, shape=(n.value,)).copy() for p in (px, py, pz)], axis=1)
    finally:
        for p in (px, py, pz):
            L.gpr_free(p)
    return out


class Context:
    def __init__(self, devices=None):
        self._h = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            _check(lib().gpr_ctx_create(arr, len(devices), C.byref(self._h)))
        else:
            _check(lib().gpr_ctx_create(None, 0, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().gpr_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_devices(self):
        return lib().gpr_ctx_num_devices(self._h)

    def set_fit_peers(self, handles, n):
        """handles: list of 128-byte CUDA IPC handle blobs (Model.ipc_export of the other ranks' replicas)."""
        blob = b"".join(handles)
        buf = C.create_string_buffer(blob, len(blob)) if blob else None
        _check(lib().gpr_ctx_set_fit_peers(self._h, C.cast(buf, C.c_void_p) if buf else None, len(handles), int(n)))

    def clear_fit_peers(self):
        _check(lib().gpr_ctx_clear_fit_peers(self._h))

    @property
    def last_fit_published(self):
        return bool(lib().gpr_ctx_last_fit_published(self._h))

    def timings(self):
        t = Timings()
        _check(lib().gpr_last_timings(self._h, C.byref(t)))
        return t.as_dict()


class Model:
    """Handle on a fitted model (gp_regression::Model, gp_regressor.hpp:71-87)."""

    def __init__(self, ctx, handle, with_normals=False):
        self.ctx, self._h, self.with_normals = ctx, handle, with_normals

    def close(self):
        if self._h:
            lib().gpr_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n(self):
        return lib().gpr_model_size(self._h)

    @property
    def n_tail(self):
        """Points eliminated as the dense indefinite pivot block (0 for an SPD covariance matrix)."""
        return lib().gpr_model_tail_size(self._h)

    def get(self):
        n = self.n
        alpha, R = np.zeros(n), C.c_double()
        N = np.zeros((n, 3), order="F") if self.with_normals else None
        _check(lib().gpr_model_get(self._h, _p(alpha), C.byref(R), _p(N)))
        return {"alpha": alpha, "R": R.value, "normals": None if N is None else np.ascontiguousarray(N)}

    @property
    def alpha(self):
        return self.get()["alpha"]

    @property
    def R(self):
        return self.get()["R"]

    def factor(self):
        n = self.n
        Lm = np.zeros((n, n), order="F")
        _check(lib().gpr_model_get_factor(self._h, _p(Lm)))
        return Lm

    def ipc_export(self):
        """128-byte CUDA IPC handle blob of this replica's factor buffers (L, Dinv) for Context.set_fit_peers."""
        buf = C.create_string_buffer(128)
        _check(lib().gpr_model_ipc_export(self.ctx._h, self._h, C.cast(buf, C.c_void_p)))
        return buf.raw

    def state(self, with_linv=False):
        s = ModelState()
        _check(lib().gpr_model_state_get(self.ctx._h, self._h, int(with_linv), C.byref(s)))
        return s


class GPRegressor:
    """Python mirror of gp_regression::GPRegressor<CovType> (gp_regressor.hpp:92-574) over the C-ABI.

    kernel: 'thin_plate' (R), 'gaussian' (sigma, length) or 'laplace' (sigma, length); the defaults
    are the reference's default constructors (R=1; sigma=length=1)."""

    def __init__(self, kernel="thin_plate", p0=1.0, p1=1.0, ctx=None):
        self.ctx = ctx or Context()
        self.set_cov_function(kernel, p0, p1)

    def set_cov_function(self, kernel, p0=1.0, p1=1.0):          # setCovFunction, :488-491
        self.kernel = KernelT(KINDS[kernel], float(p0), float(p1))

    def create(self, x, y, z, label, sigma2=None, with_normals=False):   # create<withNormals>, :110-182
        if x is None:
            raise GPRegressionException("Empty data pointer")
        x, y, z, label, sigma2 = map(_arr, (x, y, z, label, sigma2))
        if len(x) == 0 and len(y) == 0 and len(z) == 0 and len(label) == 0:
            raise GPRegressionException("All input data is empty!")
        _same_length(x, y, z, label, sigma2)
        h = C.c_void_p()
        _check(lib().gpr_fit(self.ctx._h, _p(x), _p(y), _p(z), _p(label), _p(sigma2), len(x), self.kernel,
                             int(with_normals), C.byref(h)))
        return Model(self.ctx, h, with_normals)

    def evaluate(self, model, qx, qy, qz, var=False, grad=False, tangent=False, label=None):
        """The four evaluate overloads (:194, :222, :282, :332): returns f[, v[, N[, Tx, Ty]]]."""
        if model is None or not model._h:
            raise GPRegressionException("Empty Model pointer")
        if qx is None:
            raise GPRegressionException("Empty data pointer")
        if label is not None and len(label):
            raise GPRegressionException("Query is already labeled!")
        qx, qy, qz = map(_arr, (qx, qy, qz))
        q = len(qx)
        if q == 0:
            raise GPRegressionException("All input data is empty!")
        _same_length(qx, qy, qz)
        grad = grad or tangent
        f = np.zeros(q)
        v = np.zeros(q) if var else None
        g, tx, ty = (np.zeros((q, 3), order="F") if c else None for c in (grad, tangent, tangent))
        _check(lib().gpr_predict(self.ctx._h, model._h, _p(qx), _p(qy), _p(qz), q, _p(f), _p(v), _p(g), _p(tx), _p(ty)))
        out = [f]
        if var:
            out.append(v)
        if grad:
            out.append(np.ascontiguousarray(g))
        if tangent:
            out += [np.ascontiguousarray(tx), np.ascontiguousarray(ty)]
        return out[0] if len(out) == 1 else tuple(out)

    def update(self, model, x, y, z, label, sigma2=None):        # update<withNormals>, :367-479
        if model is None or not model._h:
            raise GPRegressionException("Empty model pointer")
        x, y, z, label, sigma2 = map(_arr, (x, y, z, label, sigma2))
        _same_length(x, y, z, label, sigma2)
        _check(lib().gpr_append(self.ctx._h, model._h, _p(x), _p(y), _p(z), _p(label), _p(sigma2), len(x)))

    def reserve(self, model, capacity):
        """Pre-allocate room for `capacity` points (incremental appends then never reallocate)."""
        _check(lib().gpr_model_reserve(self.ctx._h, model._h, int(capacity)))

    def sample_isosurface(self, model, lo=-1.01, hi=1.01, step=0.07, tol=0.01, var=True, capacity=None):
        """Batched counterpart of the node's fakeDeterministicSampling (src/gp_node.cpp:998-1100): lattice points
        with |f| <= tol, as (points (k,3), f, var).  Defaults are the node's (scale 1.01, sample_res 0.07, 0.01)."""
        cnt = C.c_size_t()
        auto = capacity is None
        if auto:
            na, a = 0, lo
            while a <= hi:
                na, a = na + 1, a + step
            capacity = min(na ** 3, 1 << 22)
        xs, ys, zs, f = (np.zeros(capacity) for _ in range(4))
        v = np.zeros(capacity) if var else None
        _check(lib().gpr_sample_isosurface(self.ctx._h, model._h, lo, hi, step, tol, capacity, _p(xs), _p(ys), _p(zs), _p(f), _p(v), C.byref(cnt)))
        if auto and cnt.value > capacity:          # more survivors than room: once more with the exact size
            return self.sample_isosurface(model, lo, hi, step, tol, var, cnt.value)
        k = min(cnt.value, capacity)
        return np.stack([xs[:k], ys[:k], zs[:k]], axis=1), f[:k], (None if v is None else v[:k])

    def project(self, model, points, normals, f_tol=1e-2, improve_tol=1e-7, max_iter=500, step_mul=0.001):
        """Batched AtlasBase::project (include/atlas/atlas.hpp:201-276, same defaults): gradient-descent projection of
        every row of `points` (k,3) onto f = 0, starting with the un-normalised gradients `normals` (k,3).
        Returns (projected (k,3), status (k,) ints: iterations used, or -max_iter)."""
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3).T)
        nrm = np.ascontiguousarray(np.asarray(normals, dtype=np.float64).reshape(-1, 3).T)
        k = pts.shape[1]
        out = np.zeros((3, k))
        st = np.zeros(k, dtype=np.int32)
        _check(lib().gpr_project(self.ctx._h, model._h, _p(pts[0]), _p(pts[1]), _p(pts[2]), _p(nrm[0]), _p(nrm[1]), _p(nrm[2]), k,
                                 f_tol, improve_tol, int(max_iter), step_mul, _p(out[0]), _p(out[1]), _p(out[2]),
                                 st.ctypes.data_as(C.POINTER(C.c_int))))
        return np.ascontiguousarray(out.T), st

    def sample_marching(self, model, lo=-1.1, hi=1.1, step=0.1, leaf=0.06, leaf_pass=0.02, tol=0.01, var=True, capacity=1 << 20):
        """Batched counterpart of the node's marchingSampling (src/gp_node.cpp:1103-1292; defaults of its call at :258):
        returns (points (k,3), f, var, cubes_visited)."""
        cnt, cubes = C.c_size_t(), C.c_size_t()
        xs, ys, zs, f = (np.zeros(capacity) for _ in range(4))
        v = np.zeros(capacity) if var else None
        _check(lib().gpr_sample_marching(self.ctx._h, model._h, lo, hi, step, leaf, leaf_pass, tol, capacity, _p(xs), _p(ys), _p(zs),
                                         _p(f), _p(v), C.byref(cnt), C.byref(cubes)))
        k = min(cnt.value, capacity)
        return np.stack([xs[:k], ys[:k], zs[:k]], axis=1), f[:k], (None if v is None else v[:k]), cubes.value

    def sample_charts(self, model, frames, counts, r=None, th=None, seed=0):
        """Batched AtlasVariance::sampleOnChart (include/atlas/atlas_variance.hpp:147-219).  frames: (c, 13) rows
        [C, N, Tx, Ty, R]; counts: (c,) samples per chart; r, th: optional uniform variates (sum(counts),) each.
        Returns (samples (total, 3), f, v, order) with order[o_c + k] = in-chart index of the k-th largest variance."""
        fr = np.ascontiguousarray(np.asarray(frames, dtype=np.float64).reshape(-1, 13))
        cn = np.ascontiguousarray(np.asarray(counts, dtype=np.uint64))
        total = int(cn.sum())
        r, th = _arr(r), _arr(th)
        if (r is not None and r.size != total) or (th is not None and th.size != total) or cn.size != fr.shape[0]:
            raise GPRegressionException("Inconsistent input data sizes")
        sx, sy, sz_, f, v = (np.zeros(total) for _ in range(5))
        order = np.zeros(total, dtype=np.uint64)
        _check(lib().gpr_sample_chart(self.ctx._h, model._h, _p(fr), cn.ctypes.data_as(C.POINTER(C.c_size_t)), fr.shape[0], _p(r), _p(th),
                                      int(seed), _p(sx), _p(sy), _p(sz_), _p(f), _p(v), order.ctypes.data_as(C.POINTER(C.c_size_t))))
        return np.stack([sx, sy, sz_], axis=1), f, v, order.astype(np.int64)

    def save(self, model, path, with_factor=True):
        _check(lib().gpr_model_save(self.ctx._h, model._h, str(path).encode(), int(with_factor)))

    def load(self, path, with_normals=False):
        h = C.c_void_p()
        _check(lib().gpr_model_load(self.ctx._h, str(path).encode(), C.byref(h)))
        return Model(self.ctx, h, with_normals)

    def prepare_variance(self, model):
        _check(lib().gpr_model_prepare_variance(self.ctx._h, model._h))

    def solve(self, model, B):
        """K^-1 B through the resident factor (the reference's public Model::cholesker.solve, gp_regressor.hpp:81)."""
        B = np.asarray(B, dtype=np.float64)
        Bf = np.asfortranarray(B.reshape(model.n, -1))
        X = np.zeros_like(Bf, order="F")
        _check(lib().gpr_model_solve(self.ctx._h, model._h, _p(Bf), Bf.shape[1], _p(X)))
        return np.ascontiguousarray(X).reshape(B.shape)

    def evaluate_device(self, model, d_qx, d_qy, d_qz, q, d_f, d_var=None, d_grad=None):
        """Device-pointer variant (ints from tensor.data_ptr()); no host copies."""
        _check(lib().gpr_predict_device(self.ctx._h, model._h, d_qx, d_qy, d_qz, q, d_f, d_var, d_grad))

    def create_replica(self, n, R, with_linv, n_tail=0):
        h = C.c_void_p()
        _check(lib().gpr_model_create_replica_tail(self.ctx._h, n, n_tail, self.kernel, float(R), int(with_linv), C.byref(h)))
        return Model(self.ctx, h, False)


# ---- tile-engine self-tests (thin wrappers used by tests/) ----------------------------------------
def selftest_gemm(A, B, b_kmajor=False):
    """C = A @ B.T on the DMMA tile engine.  A: (M,k), B: (N,k); M, N multiples of 128, k of 16."""
    A = np.asfortranarray(A, dtype=np.float64)
    M, k = A.shape
    Nn = B.shape[0]
    Bm = np.asfortranarray(B.T if b_kmajor else B, dtype=np.float64)    # K-major source is k x N column-major
    Cm = np.zeros((M, Nn), order="F")
    _check(lib().gpr_selftest_gemm(_p(A), _p(Bm), int(b_kmajor), _p(Cm), M // 128, Nn // 128, k))
    return np.ascontiguousarray(Cm)


def selftest_leaf(T):
    T = np.asfortranarray(T, dtype=np.float64).copy(order="F")
    inv = np.zeros((128, 128), order="F")
    info = C.c_int()
    _check(lib().gpr_selftest_leaf(_p(T), _p(inv), C.byref(info)))
    return np.ascontiguousarray(T), np.ascontiguousarray(inv), info.value


def selftest_factor(A, want_inverse=False, serial=False):
    A = np.asfortranarray(A, dtype=np.float64).copy(order="F")
    N = A.shape[0]
    X = np.zeros((N, N), order="F") if want_inverse else None
    piv = C.c_longlong()
    rc = lib().gpr_selftest_factor(_p(A), N // 128, _p(X), int(serial), C.byref(piv))
    if rc not in (GPR_OK, GPR_ERR_NOT_SPD):
        _check(rc)
    return np.tril(A), (None if X is None else np.ascontiguousarray(X)), piv.value


def selftest_i8gemm(A, B, levels=None, tri=False, skip_zero_blocks=False):
    """Raw level accumulators of the INT8 tensor-core engine: A (S, M, K), B (S, N, K) int8 -> C (levels, M, N) int32 with
    C[l] = sum_{t+u=l} A[t] @ B[u].T (row tile r of a lower-triangular A only visits k < 128 (r + 1))."""
    A = np.ascontiguousarray(A, dtype=np.int8)
    B = np.ascontiguousarray(B, dtype=np.int8)
    S, M, K = A.shape
    N = B.shape[1]
    levels = S if levels is None else levels
    Cm = np.zeros((levels, M, N), dtype=np.int32)
    _check(lib().gpr_selftest_i8gemm(A.ctypes.data, B.ctypes.data, S, levels, M, N, K, int(tri), int(skip_zero_blocks), Cm.ctypes.data))
    return Cm


def selftest_peak(which, ctas_per_sm=4):
    t = C.c_double()
    _check(lib().gpr_selftest_peak(int(which), int(ctas_per_sm), C.byref(t)))
    return t.value


def selftest_factor_trace(n_tiles):
    """(trace[ntasks,4] in ns relative to the first claim, leaf_cycles[2]) of the tile-task Cholesky."""
    nt = n_tiles * (n_tiles + 1) // 2
    tr = (C.c_longlong * (4 * nt))()
    cy = (C.c_longlong * 8)()
    _check(lib().gpr_selftest_factor_trace(n_tiles, tr, cy))
    t = np.frombuffer(tr, dtype=np.int64).reshape(nt, 4).copy()
    return t - t[:, 0].min(), np.array(list(cy))

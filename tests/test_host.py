"""CPU suite, part 3: host-side logic — workload generators, the PCD reader, the C++ drop-in headers
(compile + argument checks, no GPU), and the world_size-2 gloo run of the query-sharding plumbing."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLD, ROOT, load_golden, relerr


def test_pcd_reader_matches_committed_clouds(gpr):
    ref = "/root/reference/resources"
    if not os.path.isdir(ref):
        pytest.skip("/root/reference not present on this box")
    for name, n in (("mugD", 262), ("kettle", 697), ("jug", 432)):
        xyz = gpr.workloads.read_pcd_xyz(os.path.join(ref, name + ".pcd"))
        assert xyz.shape == (n, 3)
        assert np.array_equal(xyz.astype(np.float32), np.load(os.path.join(GOLD, name + "_xyz.npy")))


def test_pcd_reader_ascii_and_binary(gpr, tmp_path):
    pts = np.random.default_rng(0).standard_normal((17, 3)).astype(np.float32)
    head = "# .PCD v0.7\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 17\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 17\n"
    a = tmp_path / "a.pcd"
    a.write_text(head + "DATA ascii\n" + "\n".join(" ".join(repr(float(v)) for v in p) for p in pts) + "\n")
    b = tmp_path / "b.pcd"
    b.write_bytes((head + "DATA binary\n").encode() + pts.tobytes())
    assert np.allclose(gpr.workloads.read_pcd_xyz(str(a)), pts, atol=1e-7)
    assert np.array_equal(gpr.workloads.read_pcd_xyz(str(b)).astype(np.float32), pts)


def test_node_operating_point(gpr):
    W = gpr.workloads
    ext = W.external_sphere()
    assert ext.shape == (15, 3) and np.allclose(np.linalg.norm(ext, axis=1), 2.0)
    P, y, s2 = W.node_training_set(np.load(os.path.join(GOLD, "mugD_xyz.npy")))
    g = load_golden("ref_mugD_thinplate")
    assert np.array_equal(P, g["P"]) and np.array_equal(y, g["y"]) and np.array_equal(s2, g["s2"])
    assert abs(np.linalg.norm(P[:262], axis=1).max() - 1.0) < 1e-6
    assert abs(W.max_pairwise_distance(P) - g["R"]) < 1e-12
    assert W.node_grid().shape == (29 ** 3, 3)


def test_synthetic_generators(gpr):
    W = gpr.workloads
    P, y, s2 = W.synthetic_cloud(1024, seed=0)
    P2, _, _ = W.synthetic_cloud(1024, seed=0)
    assert np.array_equal(P, P2) and P.shape == (1024, 3) and y.sum() == 256 and np.all(s2 == 0.1)
    assert np.allclose(np.linalg.norm(P[:768], axis=1), 1.0) and np.allclose(np.linalg.norm(P[768:], axis=1), 2.0)
    assert W.max_pairwise_distance(P) <= 4.0 < W.SYNTH_R
    g = W.grid_slab(8, 2, 5)
    assert g.shape == (3 * 64, 3) and np.abs(g).max() <= W.SYNTH_GRID_HALF
    # the farthest query-train distance stays inside R: the thin-plate kernel remains a valid covariance
    assert 2.0 + W.SYNTH_GRID_HALF * np.sqrt(3) <= W.SYNTH_R
    full = np.vstack([W.grid_slab(8, z, z + 1) for z in range(8)])
    assert np.array_equal(full, W.grid_slab(8, 0, 8))
    b = W.touch_batches(3, 32)
    assert len(b) == 3 and b[0][0].shape == (32, 3) and np.all(b[0][2] == 0.05)


def test_publish_policy(monkeypatch):
    """The factor is pushed to the replicas from inside the fit only while that hides in the factorisation time."""
    import gpr_b200  # noqa: F401  (registers the package alias)
    from gaussian_object_modelling_b200 import distributed as D
    monkeypatch.delenv("GPR_FIT_PUBLISH", raising=False)
    assert D.publish_pays(16384, 1) and D.publish_pays(16384, 2)
    assert not D.publish_pays(16384, 3) and not D.publish_pays(16384, 7)
    assert D.publish_pays(65536, 7) and not D.publish_pays(3000, 1) and not D.publish_pays(16384, 0)
    monkeypatch.setenv("GPR_FIT_PUBLISH", "1")
    assert D.publish_pays(3000, 7)
    monkeypatch.setenv("GPR_FIT_PUBLISH", "0")
    assert not D.publish_pays(65536, 1)


def test_shard_ranges_cover_queries():
    from gaussian_object_modelling_b200 import distributed as D
    for q in (1, 7, 1000, 256 ** 3):
        for w in (1, 2, 4, 8):
            r = [D.shard_range(q, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == q and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _build_driver(tmp):
    exe = os.path.join(tmp, "dropin_driver")
    pkg = os.path.join(ROOT, "gaussian-object-modelling_b200")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "dropin_driver.cpp"), "-o", exe, "-L", pkg, "-lgpr_b200",
           "-Wl,-rpath," + pkg, "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True)
    return exe


def test_dropin_headers_compile_as_cxx11_and_check_arguments(gpr, tmp_path):
    """The reference is C++11 (CMakeLists.txt:4); the drop-in headers must build as such without Eigen,
    and reproduce the reference's exception messages before any GPU work."""
    exe = _build_driver(str(tmp_path))
    out = subprocess.run([exe, "--errors"], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "0: Empty data pointer"
    assert out[1] == "1: All input data is empty!"
    assert out[2] == "2: Empty Model pointer"
    assert out[3] == "3: Empty model pointer"
    assert out[4] == "4: Query is already labeled!"
    assert out[5] == "basis 0 0 1 1 0 0 0 1 0" and out[6] == "basis 1 0 0 0 1 0 0 0 1"      # Appendix B.6
    k = [float(v) for v in out[7].split()[1:]]
    assert k[:4] == [4.0, -6.0, 8.0, 0.0]                                                      # Appendix B.1
    assert abs(k[4] - np.exp(-1)) < 1e-6 and abs(k[5] + np.exp(-1)) < 1e-6


_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
import gpr_b200, oracle
from gaussian_object_modelling_b200 import distributed as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
W = gpr_b200.workloads
P, y, s2 = W.synthetic_cloud(256, seed=0)
Q = W.grid_slab(12, 0, 12)
# rank 0 "fits" (here: the CPU oracle stands in for the GPU fit), then the predict state is broadcast once
n = len(P)
alpha = np.zeros(n); L = np.zeros((n, n)); X = np.zeros((n, 3))
if rank == 0:
    m = oracle.blas_fit(P, y, s2, "thin_plate", W.SYNTH_R, 0.0)
    alpha[:] = m["alpha"]; L[:] = m["L"]; X[:] = P
D.broadcast_arrays([X, alpha, L])
model = dict(L=L, alpha=alpha, P=X, kind="thin_plate", p0=W.SYNTH_R, p1=0.0)
a, b = D.shard_range(len(Q), rank, 2)
f, v = oracle.blas_predict(model, Q[a:b])
F = D.gather_shards(f, len(Q), rank, 2); V = D.gather_shards(v, len(Q), rank, 2)
if rank == 0:
    f1, v1 = oracle.blas_predict(oracle.blas_fit(P, y, s2, "thin_plate", W.SYNTH_R, 0.0), Q)
    # row-blocked BLAS calls may differ in the last bit; the sharded CUDA path is bit-identical (GPU test)
    assert np.abs(F - f1).max() <= 1e-12 and np.abs(V - v1).max() <= 1e-10, (np.abs(F - f1).max(), np.abs(V - v1).max())
    print("GLOO_OK", len(Q), a, b)
dist.barrier()
dist.destroy_process_group()
"""


def test_query_sharding_world_size_2_gloo(tmp_path, orc):
    """N>1 host logic on CPU: rank 0 fits, one broadcast of (X, alpha, factor), every rank predicts its
    contiguous query block, the gathered result equals the unsharded one."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GLOO_OK" in outs[0]


def test_standalone_cmake_build(tmp_path):
    """the library builds standalone, without ROS / catkin (the reference is a catkin package): configure and
    build CMakeLists.txt (nvcc cross-compiles sm_100a without a GPU), then run the drop-in driver's argument checks."""
    import shutil
    if shutil.which("cmake") is None or not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("cmake or nvcc not available")
    b = str(tmp_path / "build")
    cfg = subprocess.run(["cmake", "-S", ROOT, "-B", b, "-DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc"], capture_output=True, text=True)
    assert cfg.returncode == 0, cfg.stdout[-2000:] + cfg.stderr[-2000:]
    bld = subprocess.run(["cmake", "--build", b, "-j8"], capture_output=True, text=True)
    assert bld.returncode == 0, bld.stdout[-2000:] + bld.stderr[-2000:]
    out = subprocess.run([os.path.join(b, "dropin_driver"), "--errors"], capture_output=True, text=True)
    assert out.returncode == 0 and "All input data is empty!" in out.stdout
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(b, "libgpr_b200.so")], capture_output=True, text=True).stdout
    assert " gpr_fit" in syms and " gpr_predict" in syms and " gpr_append" in syms


def _lzf_compress(data):
    """Greedy LZF compressor (test helper): emits literal runs and back references, like PCL's writer does."""
    out, lit, i, n = bytearray(), bytearray(), 0, len(data)
    table = {}

    def flush():
        j = 0
        while j < len(lit):
            run = lit[j:j + 32]
            out.append(len(run) - 1)
            out.extend(run)
            j += 32
        lit.clear()

    while i < n:
        key = bytes(data[i:i + 3])
        ref = table.get(key) if len(key) == 3 else None
        if len(key) == 3:
            table[key] = i
        if ref is not None and 0 < i - ref <= 8192:
            length = 3
            while i + length < n and length < 264 and data[ref + length] == data[i + length]:
                length += 1
            flush()
            dist, l2 = i - ref - 1, length - 2
            if l2 < 7:
                out.append((l2 << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8))
                out.append(l2 - 7)
            out.append(dist & 0xFF)
            i += length
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def test_c_abi_pcd_reader(gpr, tmp_path):
    """gpr_pcd_read_xyz (csrc/gpr_io.cu; the reference uses PCL, src/gp_node.cpp:557): ascii / binary / binary_compressed,
    x y z among other fields, float32 widened to double; malformed input is refused with a message.  No GPU needed."""
    import struct
    rng = np.random.default_rng(5)
    n = 300
    xyz = rng.standard_normal((n, 3)).astype(np.float32)
    xyz[40:90] = xyz[40]                                       # repeated records: the LZF stream gets back references
    rgba = rng.integers(0, 1 << 32, size=n, dtype=np.uint32)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS rgba x y z\nSIZE 4 4 4 4\nTYPE U F F F\n"
           "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n")
    rec = np.zeros(n, dtype=[("rgba", "<u4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4")])
    rec["rgba"], rec["x"], rec["y"], rec["z"] = rgba, xyz[:, 0], xyz[:, 1], xyz[:, 2]
    (tmp_path / "b.pcd").write_bytes((hdr % (n, n, "binary")).encode() + rec.tobytes())
    soa = rgba.tobytes() + xyz[:, 0].tobytes() + xyz[:, 1].tobytes() + xyz[:, 2].tobytes()
    comp = _lzf_compress(soa)
    assert len(comp) < len(soa)                                # back references were emitted
    (tmp_path / "c.pcd").write_bytes((hdr % (n, n, "binary_compressed")).encode() + struct.pack("<II", len(comp), len(soa)) + comp)
    with open(tmp_path / "a.pcd", "w") as fh:
        fh.write(hdr % (n, n, "ascii"))
        for i in range(n):
            fh.write("%d %.9g %.9g %.9g\n" % (rgba[i], xyz[i, 0], xyz[i, 1], xyz[i, 2]))
    for name in ("a.pcd", "b.pcd", "c.pcd"):
        got = gpr.pcd_read_xyz(tmp_path / name)
        assert got.dtype == np.float64 and np.array_equal(got, xyz.astype(np.float64)), name
        assert np.array_equal(got, gpr.workloads.read_pcd_xyz(str(tmp_path / name)))       # the Python reader of the benches
    # the C++ header's loadPCD (drop-in side of the same reader)
    exe = _build_driver(str(tmp_path))
    out = subprocess.run([exe, "--pcd", str(tmp_path / "c.pcd")], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "pcd" and int(out[1]) == n and int(out[5]) == 0
    assert np.allclose([float(v) for v in out[2:5]], xyz.astype(np.float64).sum(axis=0), rtol=1e-12, atol=1e-12)
    out = subprocess.run([exe, "--pcd", str(tmp_path / "missing.pcd")], capture_output=True, text=True, check=True).stdout
    assert out.startswith("exception") and "cannot open" in out
    (tmp_path / "bad.pcd").write_bytes((hdr % (n, n, "binary_compressed")).encode() + struct.pack("<II", len(comp), len(soa)) + comp[:-7])
    (tmp_path / "short.pcd").write_bytes((hdr % (n, n, "binary")).encode() + rec.tobytes()[:-5])
    (tmp_path / "nofields.pcd").write_bytes(b"VERSION 0.7\nDATA ascii\n1 2 3\n")
    for name in ("bad.pcd", "short.pcd", "nofields.pcd", "missing.pcd"):
        with pytest.raises(gpr.GPRegressionException):
            gpr.pcd_read_xyz(tmp_path / name)
    # the reference's own clouds, where they exist (not on the GPU box), against the committed decoded fixtures
    res = "/root/reference/resources"
    if os.path.isdir(res):
        for name in ("mugD", "kettle", "jug"):
            gold = np.load(os.path.join(ROOT, "tests", "golden", name + "_xyz.npy")).astype(np.float64)
            assert np.array_equal(gpr.pcd_read_xyz(os.path.join(res, name + ".pcd")), gold)


def test_digit_slicing_scheme_on_the_cpu():
    """The integer-slicing arithmetic of the INT8 kernels (csrc/gpr_ozaki.cu), emulated with numpy: (i) the rounding
    without conversion instructions, v + 1.5*2^52 - 1.5*2^52, equals rint(v) and leaves the integer in the low word of
    the sum; (ii) S base-254 digits reconstruct a scaled value to 0.5 / (127 * 254^(S-1)); (iii) a left-looking Cholesky
    whose updates use the sliced factor (7 digits, per-row power-of-two scales, levels t + u < 7 only, exact integer
    products) is as accurate as the plain float64 one — the scheme of launch_cholesky_int8 at a size numpy handles."""
    rng = np.random.default_rng(11)
    v = np.concatenate([rng.uniform(-127.5, 127.5, 4000), np.arange(-127, 128) + 0.5, np.arange(-127, 128) - 0.5])
    MAGIC = 6755399441055744.0
    m = v + MAGIC
    assert np.array_equal(m - MAGIC, np.rint(v))
    low = (m.view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32)
    assert np.array_equal(low.astype(np.float64), np.rint(v))

    def digits(x, S, base=254.0):                       # x in [-1, 1]
        r, out = x * (base / 2.0), []
        for _ in range(S):
            it = np.rint(r)
            out.append(it.astype(np.int64))
            r = (r - it) * base
        return out

    x = rng.uniform(-1.0, 1.0, 2000)
    for S in (6, 7):
        d = digits(x, S)
        assert max(int(np.abs(t).max()) for t in d) <= 127
        rec = sum(t.astype(np.longdouble) / (np.longdouble(127.0) * np.longdouble(254.0) ** i) for i, t in enumerate(d))
        assert float(np.abs(rec - x.astype(np.longdouble)).max()) <= 0.5 / (127.0 * 254.0 ** (S - 1)) + 3e-16   # + the rounding of the digit recursion itself

    n, P, S = 192, 48, 7
    pts = rng.random((n, 3))
    K = np.exp(-((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1) / 0.5) + 1e-5 * np.eye(n)
    dscale = 10.0 ** rng.uniform(-2.0, 2.0, n)
    A = K * dscale[:, None] * dscale[None, :]
    s = 2.0 ** (np.floor(np.log2(np.sqrt(A.diagonal()))) + 1)          # power of two above sqrt(K_ii)
    L = np.zeros((n, n))
    W = A.copy()
    for c0 in range(0, n, P):
        c1 = min(n, c0 + P)
        if c0:
            D = digits(L[c0:, :c0] / s[c0:, None], S)                  # slices of the finished columns, rows from c0 on
            upd = np.zeros((n - c0, c1 - c0))
            for l in range(S):                                         # levels t + u = l < S, exact int64 products
                acc = sum(D[t] @ D[l - t][:c1 - c0].T for t in range(l + 1))
                upd += acc.astype(np.float64) / (127.0 ** 2 * 254.0 ** l)
            W[c0:, c0:c1] -= upd * s[c0:, None] * s[None, c0:c1]
        for j in range(c0, c1):                                        # the panel itself in float64, k restricted to the panel
            W[j, j] = np.sqrt(W[j, j] - L[j, c0:j] @ L[j, c0:j])
            L[j, j] = W[j, j]
            L[j + 1:, j] = (W[j + 1:, j] - L[j + 1:, c0:j] @ L[j, c0:j]) / L[j, j]
    L64 = np.linalg.cholesky(A)
    err = lambda F: np.abs((F.astype(np.longdouble) @ F.T.astype(np.longdouble) - A) / (dscale[:, None] * dscale[None, :])).max()
    assert err(L) <= 3.0 * err(L64) + 1e-16

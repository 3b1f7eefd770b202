"""Inputs for the five BASELINE.json configs: PCD decoding, the node's preprocessing
recipe and the synthetic sphere/ellipsoid generator.  Host-side numpy only; no GPU, no oracle.

Reference anchors (paths relative to /root/reference):
  * demean + scale to the unit ball ............ src/gp_node.cpp:85-117
  * 15 external points on the r=2 sphere, y=+1 . src/gp_node.cpp:793-850 (ang_div=5, lin_div=3)
  * surface points y=0, sigma2=0.1 everywhere .. src/gp_node.cpp:853-888, :16
  * training order = cloud points then external  src/gp_node.cpp:898-914
  * R = max pairwise distance ................... tests/test_gp.cpp:125-132
  * sampling grid [-1.01,1.01]^3 step 0.07 ...... src/gp_node.cpp:1025-1030
"""
import math
import struct

import numpy as np


# ------------------------------------------------------------------------------------------
# Minimal PCD v0.7 reader: ascii / binary / binary_compressed (LZF), float32 fields -> float64 xyz.
# ------------------------------------------------------------------------------------------
def _lzf_decompress(src, out_len):
    out = bytearray(out_len)
    i = o = 0
    n = len(src)
    while i < n:
        ctrl = src[i]
        i += 1
        if ctrl < 32:                       # literal run
            run = ctrl + 1
            out[o:o + run] = src[i:i + run]
            i += run
            o += run
        else:                               # back reference
            length = ctrl >> 5
            if length == 7:
                length += src[i]
                i += 1
            ref = o - ((ctrl & 0x1F) << 8) - src[i] - 1
            i += 1
            for _ in range(length + 2):     # may overlap: byte by byte
                out[o] = out[ref]
                o += 1
                ref += 1
    if o != out_len:
        raise ValueError("LZF stream ended at %d of %d bytes" % (o, out_len))
    return bytes(out)


def read_pcd_xyz(path):
    """Return the x,y,z fields of a PCD file as an (n,3) float64 array."""
    with open(path, "rb") as fh:
        raw = fh.read()
    header, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, val = line.partition(" ")
        header[key] = val.split()
        if key == "DATA":
            break
    fields, sizes = header["FIELDS"], [int(s) for s in header["SIZE"]]
    counts = [int(c) for c in header.get("COUNT", ["1"] * len(fields))]
    npts = int(header["POINTS"][0])
    mode = header["DATA"][0]
    widths = [s * c for s, c in zip(sizes, counts)]
    offs = np.concatenate([[0], np.cumsum(widths)])
    idx = [fields.index(a) for a in "xyz"]
    if mode == "ascii":
        rows = np.array([ln.split() for ln in raw[pos:].decode().strip().splitlines()], dtype=object)
        return np.stack([rows[:, i].astype(np.float32) for i in idx], axis=1).astype(np.float64)
    if mode == "binary":
        rec = np.frombuffer(raw, dtype=np.uint8, count=npts * int(offs[-1]), offset=pos).reshape(npts, -1)
        cols = [rec[:, offs[i]:offs[i] + 4].copy().view(np.float32)[:, 0] for i in idx]
        return np.stack(cols, axis=1).astype(np.float64)
    if mode == "binary_compressed":
        csize, usize = struct.unpack_from("<II", raw, pos)
        buf = _lzf_decompress(raw[pos + 8:pos + 8 + csize], usize)
        cols = []                            # decompressed layout is SoA: all of field 0, then field 1, ...
        for i in idx:
            start = int(offs[i]) * npts
            cols.append(np.frombuffer(buf, dtype=np.float32, count=npts, offset=start))
        return np.stack(cols, axis=1).astype(np.float64)
    raise ValueError("unsupported PCD DATA mode %r" % mode)


# ------------------------------------------------------------------------------------------
# The node's operating point.
# ------------------------------------------------------------------------------------------
def demean_and_normalize(xyz):
    """float32 centroid subtraction and scaling by the largest norm (src/gp_node.cpp:85-117)."""
    p = np.asarray(xyz, dtype=np.float32)
    p = p - p.mean(axis=0, dtype=np.float32)
    scale = float(np.sqrt((p.astype(np.float64) ** 2).sum(axis=1)).max())
    return (p * np.float32(1.0 / scale)).astype(np.float64)


def external_sphere(radius=2.0, ang_div=5, lin_div=3):
    """The 15 external points of prepareExtData (src/gp_node.cpp:821-849), same loop order."""
    ang_step = math.pi * 2 / ang_div
    lin_step = 2 * radius / lin_div
    pts = []
    lin = -radius + lin_step / 2
    while lin < radius:
        ang = 0.0
        while ang < 2 * math.pi:
            r = math.sqrt(radius ** 2 - lin * lin)
            pts.append((r * math.cos(ang), r * math.sin(ang), lin))
            ang += ang_step
        lin += lin_step
    return np.array(pts, dtype=np.float64)


def max_pairwise_distance(P, block=2048):
    P = np.asarray(P, dtype=np.float64)
    best = 0.0
    for i in range(0, len(P), block):
        d = P[i:i + block, None, :] - P[None, :, :]
        best = max(best, float(np.sqrt((d * d).sum(-1)).max()))
    return best


def node_training_set(cloud_xyz, sigma2=0.1, radius=2.0):
    """cloud (label 0) then external sphere (label +1), sigma2 on every point (src/gp_node.cpp:898-914)."""
    surf = demean_and_normalize(cloud_xyz)
    ext = external_sphere(radius)
    P = np.vstack([surf, ext])
    y = np.concatenate([np.zeros(len(surf)), np.ones(len(ext))])
    return P, y, np.full(len(P), sigma2)


def node_grid(limit=1.01, step=0.07):
    """The fakeDeterministicSampling lattice (src/gp_node.cpp:1025-1030): 29^3 points."""
    axis = []
    v = -limit
    while v <= limit:
        axis.append(v)
        v += step
    a = np.array(axis)
    X, Y, Z = np.meshgrid(a, a, a, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)


# ------------------------------------------------------------------------------------------
# Synthetic clouds for configs 3-5: 3/4 of the points uniform on the unit sphere
# (or an ellipsoid) with label 0, 1/4 uniform on the r=2 sphere with label +1, sigma2 = 0.1.
# ThinPlate R = 4.2 >= 2 + 1.2*sqrt(3) = 4.08, so the kernel stays a valid covariance for every
# train-train pair (<= 4) and every query of the [-1.2,1.2]^3 grid.
# ------------------------------------------------------------------------------------------
SYNTH_R = 4.2
SYNTH_GRID_HALF = 1.2


def _unit_sphere(rng, n):
    v = rng.standard_normal((n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def synthetic_cloud(n, seed=0, semi_axes=(1.0, 1.0, 1.0), sigma2=0.1, outer_radius=2.0):
    rng = np.random.default_rng(seed)
    n_ext = n // 4
    n_surf = n - n_ext
    surf = _unit_sphere(rng, n_surf) * np.asarray(semi_axes)
    ext = _unit_sphere(rng, n_ext) * outer_radius
    P = np.vstack([surf, ext])
    y = np.concatenate([np.zeros(n_surf), np.ones(n_ext)])
    return P, y, np.full(n, sigma2)


def grid_slab(res, z0, z1, half=SYNTH_GRID_HALF):
    """Rows z0..z1-1 (z-major slabs) of the res^3 lattice on [-half,half]^3; returns (q,3).
    Query index = (iz*res + iy)*res + ix, so contiguous index ranges are z-slabs."""
    a = np.linspace(-half, half, res)
    Z, Y, X = np.meshgrid(a[z0:z1], a, a, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)


def touch_batches(n_batches, batch, seed=1, semi_axes=(1.0, 1.0, 1.0)):
    """Config 4: new surface points inside the existing hull (label 0, sigma2 = 0.05 per the touch
    noise at src/gp_node.cpp:693), so the model's R is unchanged."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_batches):
        p = _unit_sphere(rng, batch) * np.asarray(semi_axes)
        out.append((p, np.zeros(batch), np.full(batch, 0.05)))
    return out

"""Generates tests/golden/*.npz by running the reference's OWN header
(/root/reference/include/gp_regression/gp_regressor.hpp, compiled unmodified against oracle/eigen_shim
into oracle/_ref/libgpr_ref.so) on inputs derived from the reference's resources/*.pcd.

Run here, where /root/reference exists:   python oracle/make_golden.py
The GPU box has no /root/reference; tests there read the committed fixtures only.
The decoded point clouds (tests/golden/*_xyz.npy, float32 as in the PCD files) were written by the same
reader (gaussian-object-modelling_b200/workloads.py:read_pcd_xyz).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle                                    # noqa: E402
import gpr_b200                                  # noqa: E402  (host-side workloads only; no GPU use)

W = gpr_b200.workloads
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/resources"


def cloud(name):
    xyz = W.read_pcd_xyz(os.path.join(REF, name + ".pcd"))
    np.save(os.path.join(GOLD, name + "_xyz.npy"), xyz.astype(np.float32))
    return xyz


def run_case(tag, P, y, s2, kind, p0, p1, Q, normals, upd=None):
    ref = oracle.Reference(kind, p0, p1).fit(P[:, 0], P[:, 1], P[:, 2], y, s2, with_normals=normals)
    got = ref.get(K=True)
    f1, = ref.evaluate(Q[:, 0], Q[:, 1], Q[:, 2], mode=1)
    f4, v4, N4, Tx, Ty = ref.evaluate(Q[:, 0], Q[:, 1], Q[:, 2], mode=4)
    out = dict(P=P, y=y, s2=s2, Q=Q, kind=kind, p0=p0, p1=p1, R=ref.R, alpha=got["alpha"], f=f4, f_mean_only=f1, v=v4,
               grad=N4, Tx=Tx, Ty=Ty, K_row0=got["K"][0].copy(), K_diag=np.diag(got["K"]).copy(),
               K_fro=np.linalg.norm(got["K"]))
    if normals:
        out["normals"] = got["normals"]
    if upd is not None:
        Pu, yu, su = upd
        ref.update(Pu[:, 0], Pu[:, 1], Pu[:, 2], yu, su)
        out.update(Pu=Pu, yu=yu, su=su, alpha_updated=ref.get()["alpha"], R_updated=ref.R,
                   f_updated=ref.evaluate(Q[:, 0], Q[:, 1], Q[:, 2], mode=1)[0])
    np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **out)
    print(tag, "n=%d q=%d R=%.6f |alpha|max=%.3e" % (len(P), len(Q), ref.R, np.abs(got["alpha"]).max()))


def node_cases():
    """The ROS node's REAL operating point: ThinPlate(2.0) (src/gp_node.cpp:919) with the 15 external
    points at r = 2 (:16, :821-849).  K is indefinite (3 negative eigenvalues); the reference's pivoted LDLT
    handles it, a plain Cholesky cannot."""
    grid = W.node_grid()
    for name, step in (("mugD", 97), ("jug", 131)):
        P, y, s2 = W.node_training_set(W.read_pcd_xyz(os.path.join(REF, name + ".pcd")))
        run_case("ref_%s_thinplate_R2_node" % name, P, y, s2, "thin_plate", 2.0, 0.0, grid[::step], normals=True)


def lattice_cases():
    """The node's fakeDeterministicSampling lattice (src/gp_node.cpp:998-1100: 29^3 points, one evaluate(q = 1) per point,
    keep |f| <= 0.01, the variance is the intensity) run through the reference's own evaluate() on mugD, for the SPD
    setting (R = max pairwise distance) and for the node's own ThinPlate(2.0).  Stored: the lattice indices with
    |f| <= 0.02 (the kept shell plus a margin), their f and v, and the number of kept points."""
    grid = W.node_grid()
    P, y, s2 = W.node_training_set(W.read_pcd_xyz(os.path.join(REF, "mugD.pcd")))
    for tag, R in (("spd", W.max_pairwise_distance(P)), ("R2_node", 2.0)):
        ref = oracle.Reference("thin_plate", R, 0.0).fit(P[:, 0], P[:, 1], P[:, 2], y, s2)
        f, v = ref.evaluate_mt(grid[:, 0], grid[:, 1], grid[:, 2], var=True, threads=os.cpu_count() or 1, per_call=1)
        band = np.flatnonzero(np.abs(f) <= 0.02)
        np.savez_compressed(os.path.join(GOLD, "ref_mugD_lattice_%s.npz" % tag), P=P, y=y, s2=s2, p0=R, idx=band.astype(np.int32),
                            f=f[band], v=v[band], kept=int((np.abs(f) <= 0.01).sum()), lattice=len(grid))
        print("lattice", tag, "kept", int((np.abs(f) <= 0.01).sum()), "band", len(band))


def main():
    oracle.build()
    grid = W.node_grid()
    # config 1: mugD, node preprocessing, ThinPlate(R = max pairwise distance) — tests/test_gp.cpp:125-132
    P, y, s2 = W.node_training_set(cloud("mugD"))
    R = W.max_pairwise_distance(P)
    run_case("ref_mugD_thinplate", P, y, s2, "thin_plate", R, 0.0, grid[::97], normals=True,
             upd=(np.array([[0.3, 0.1, -0.2], [0.0, 0.5, 0.4], [-0.6, 0.2, 0.1]]), np.zeros(3), np.full(3, 0.05)))
    # config 2: kettle / jug, Gaussian(1,1) defaults, outputs at the training points
    P, y, s2 = W.node_training_set(cloud("kettle"))
    run_case("ref_kettle_gaussian", P, y, s2, "gaussian", 1.0, 1.0, P, normals=True)      # ALL training points (config 2)
    P, y, s2 = W.node_training_set(cloud("jug"))
    run_case("ref_jug_gaussian", P, y, s2, "gaussian", 1.0, 1.0, P, normals=True)
    run_case("ref_jug_laplace", P, y, s2, "laplace", 1.0, 1.0, grid[::211], normals=False)
    node_cases()
    lattice_cases()
    # the reference's argument checks
    msgs = [oracle.Reference().error_message(i) for i in range(4)]
    np.savez(os.path.join(GOLD, "ref_error_messages.npz"), messages=np.array(msgs))
    print(msgs)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "node":
        oracle.build()
        node_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "lattice":
        oracle.build()
        lattice_cases()
    else:
        main()

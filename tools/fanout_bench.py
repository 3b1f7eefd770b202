#!/usr/bin/env python
"""The unchanged-caller measurement: the node's thread-per-lattice-point sampling loop
(src/gp_node.cpp:1025-1038 + :1067-1100; tests/cpp/fanout_bench.cpp) — 29 slabs x 841 std::threads x one
evaluate(gp, q, f, v) with a single query — run through

  * this repository's drop-in headers over libgpr_b200.so (GPU arm: concurrent single-query calls are combined by
    the micro-batcher inside gpr_predict), and
  * the reference's own header (CPU arm: oracle/_ref/fanout_ref, the SAME source file compiled against
    /root/reference/include where that exists),

on the same box, same inputs: mugD (n = 277) and kettle (n = 712) with the node's preprocessing, for the SPD setting
(R = max pairwise distance) and the node's own ThinPlate(2.0).  Prints one JSON object; results of the two arms are
compared point by point.

  python tools/fanout_bench.py [--out gpurun_out/fanout.json] [--cases mugD:node,kettle:spd,...]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "gaussian-object-modelling_b200")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "fanout_ref")


def build_ours(tmp):
    exe = os.path.join(tmp, "fanout_ours")
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-Werror", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "fanout_bench.cpp"), "-o", exe, "-L", PKG, "-lgpr_b200",
           "-Wl,-rpath," + PKG, "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True)
    return exe


def write_input(path, P, y, s2, R):
    with open(path, "w") as fh:
        fh.write("%r\n%d\n" % (float(R), len(P)))
        for p, l, s in zip(P, y, s2):
            fh.write(" ".join(repr(float(t)) for t in (p[0], p[1], p[2], l, s)) + "\n")


def read_out(path):
    raw = np.fromfile(path, dtype=np.uint8)
    cnt = int(np.frombuffer(raw[:8], dtype=np.int64)[0])
    return (np.frombuffer(raw[8:8 + 8 * cnt], dtype=np.float64).copy(),
            np.frombuffer(raw[8 + 8 * cnt:8 + 16 * cnt], dtype=np.float64).copy())


def run_arm(exe, inp, out, env=None):
    r = subprocess.run([exe, inp, out], capture_output=True, text=True, timeout=1800, env=env)
    if r.returncode != 0:
        raise RuntimeError("%s failed: %s" % (exe, r.stderr[-2000:]))
    return json.loads(r.stdout.strip().splitlines()[-1])


def training_set(name):
    import gpr_b200
    W = gpr_b200.workloads
    xyz = np.load(os.path.join(ROOT, "tests", "golden", name + "_xyz.npy")).astype(np.float64)
    return W.node_training_set(xyz), W


def run_case(name, setting, ours_exe, tmp, with_ref=True, also_unbatched=False):
    (P, y, s2), W = training_set(name)
    R = 2.0 if setting == "node" else W.max_pairwise_distance(P)
    inp = os.path.join(tmp, "%s_%s.txt" % (name, setting))
    write_input(inp, P, y, s2, R)
    res = {"cloud": name, "setting": "ThinPlate(2.0), indefinite K (the node's own)" if setting == "node" else "ThinPlate(R = max pairwise distance), SPD",
           "n": len(P)}
    res["ours"] = run_arm(ours_exe, inp, os.path.join(tmp, "ours.bin"))
    fo, vo = read_out(os.path.join(tmp, "ours.bin"))
    if also_unbatched:
        env = dict(os.environ, GPR_MICROBATCH="0")
        res["ours_without_microbatcher"] = run_arm(ours_exe, inp, os.path.join(tmp, "ours_nb.bin"), env)
    if with_ref and os.path.exists(REF_EXE):
        res["reference"] = run_arm(REF_EXE, inp, os.path.join(tmp, "ref.bin"))
        fr, vr = read_out(os.path.join(tmp, "ref.bin"))
        big = np.abs(fr) > 1e-9
        res["parity"] = {"mean_rel_inf": float(np.abs(fo - fr).max() / np.abs(fr).max()),
                         "var_rel_inf": float(np.abs(vo - vr).max() / np.abs(vr).max()),
                         "sign_mismatches": int((np.sign(fo) != np.sign(fr))[big].sum()),
                         "kept_equal": bool(res["ours"]["kept"] == res["reference"]["kept"])}
        res["speedup_vs_reference"] = res["reference"]["total_s"] / res["ours"]["total_s"]
    return res, (fo, vo)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--cases", default="mugD:node,mugD:spd,kettle:node,kettle:spd")
    ap.add_argument("--unbatched", action="store_true", help="also time the GPU arm with GPR_MICROBATCH=0")
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        exe = build_ours(tmp)
        results = []
        for case in args.cases.split(","):
            name, setting = case.split(":")
            res, _ = run_case(name, setting, exe, tmp, also_unbatched=args.unbatched)
            results.append(res)
            print(json.dumps(res), file=sys.stderr)
    line = {"what": "node-style thread fan-out, one evaluate(q=1) per lattice point (tests/cpp/fanout_bench.cpp)", "cases": results}
    print(json.dumps(line))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(line, fh, indent=1)


if __name__ == "__main__":
    main()

"""The node's fan-out (src/gp_node.cpp:1027-1038): 841 concurrent host threads, one evaluate(q = 1, mean + variance) each,
on one shared model — for an SPD model (fused single-launch kernel) and for the node's indefinite setting (tail block).
Prints the wall time per slab of 841 calls and checks every result against one batched call."""
import os
import sys
import threading
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g

W = g.workloads
ctx = g.Context()
cloud = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mugD_xyz.npy")).astype(np.float64)
P, y, s2 = W.node_training_set(cloud)
G = W.node_grid()
for name, R in (("SPD, ThinPlate(R = max distance)", W.max_pairwise_distance(P)), ("node setting, ThinPlate(2.0), indefinite", 2.0)):
    reg = g.GPRegressor("thin_plate", R, ctx=ctx)
    m = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    reg.prepare_variance(m)
    slab = G[:841]                                   # one x-slab of the 29^3 lattice = 841 points
    fb, vb = reg.evaluate(m, slab[:, 0], slab[:, 1], slab[:, 2], var=True)
    res = [None] * 841

    def work(i):
        res[i] = reg.evaluate(m, slab[i:i + 1, 0], slab[i:i + 1, 1], slab[i:i + 1, 2], var=True)

    for rep in range(3):
        th = [threading.Thread(target=work, args=(i,)) for i in range(841)]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        wall = time.perf_counter() - t0
    f = np.array([r[0][0] for r in res]); v = np.array([r[1][0] for r in res])
    ef, ev = np.abs(f - fb).max() / np.abs(fb).max(), np.abs(v - vb).max() / np.abs(vb).max()
    t0 = time.perf_counter()
    for i in range(200):
        work(i)
    seq = (time.perf_counter() - t0) / 200
    print("%s: n=%d tail=%d | 841 threads x 1 query: %.1f ms per slab (%.1f us per call amortised), sequential %.1f us per call, "
          "one batched call of 841: see config1 | max rel diff vs batched: mean %.1e var %.1e" % (name, m.n, m.n_tail, 1e3 * wall, 1e6 * wall / 841, 1e6 * seq, ef, ev), flush=True)
    assert ef <= 1e-9 and ev <= 1e-7
print("NODE_THREADS_OK")

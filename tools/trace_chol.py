"""Timeline analysis of the tile-task Cholesky (development tool, GPU box)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpr_b200 as g

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tr, cy = g.selftest_factor_trace(nb)
print("leaf cycles potrf128=%d trinv128=%d | potrf phases: A=%d A-wait=%d B=%d C1=%d C2(warp1)=%d B(thread127)=%d" % tuple(cy))
def order(nb):          # same order as chol_task_decode (gpr_factor.cu)
    yield (0, 0)
    for j in range(nb - 1):
        yield (j + 1, j)
        yield (j + 1, j + 1)
        for i in range(j + 2, nb):
            yield (i, j)
tasks = np.array([(i, j, *tr[t]) for t, (i, j) in enumerate(order(nb))])
total = tr[:, 3].max()
print("nb=%d total %.3f ms  (%.2f TF/s)" % (nb, total / 1e6, (128 * nb) ** 3 / 3 / (total * 1e-9) / 1e12))
diag = tasks[tasks[:, 0] == tasks[:, 1]]
diag = diag[np.argsort(diag[:, 1])]
off1 = tasks[tasks[:, 0] == tasks[:, 1] + 1]
off1 = off1[np.argsort(off1[:, 1])]
for j in sorted(set([0, 1, 2, 3, nb // 4, nb // 2, 3 * nb // 4, nb - 2, nb - 1])):
    d = diag[j]
    line = "col %3d diag: claim@%9.1fus acc %7.1f leaf %7.1f pub %5.1f" % (j, d[2] / 1e3, (d[3] - d[2]) / 1e3, (d[4] - d[3]) / 1e3, (d[5] - d[4]) / 1e3)
    if j < nb - 1:
        o = off1[j]
        line += " | (j+1,j): claim@%9.1f acc_done@%9.1f solved@%9.1f (trsm+wait %6.1f) " % (o[2] / 1e3, o[3] / 1e3, o[4] / 1e3, (o[4] - o[3]) / 1e3)
    print(line)
dd = np.diff(diag[:, 5]) / 1e3
print("diag-to-diag publish interval (us): first8", np.round(dd[:8], 1), "mid", np.round(dd[nb // 2 - 2: nb // 2 + 2], 1), "last8", np.round(dd[-8:], 1))
print("mean leaf (acc->solved) on diag tasks: %.1f us" % ((diag[:, 4] - diag[:, 3]).mean() / 1e3))
busy = (tasks[:, 5] - tasks[:, 2]).sum()
wait_d = np.array([t[4] - t[3] for t in tasks if t[0] != t[1]])
print("off-diagonal tasks: mean (acc done -> solved) = %.1f us (trsm + wait for Dinv)" % (wait_d.mean() / 1e3))
print("sum of task durations / (148 * total) = %.3f" % (busy / (148 * total)))
acc_time = (tasks[:, 3] - tasks[:, 2]).sum()
ksteps = (tasks[:, 1]).sum()
print("accumulation: %.2f us per 128-wide k-step (incl. waits)" % (acc_time / max(ksteps, 1) / 1e3))

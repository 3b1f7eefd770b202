"""Turn an ncu report (--set full) into the markdown summary committed under profiles/.
  python tools/ncu_summary.py gpurun_out/prof_r1x.ncu-rep profiles/ncu_summary_r1x.md [n_train] [queries]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
q = int(sys.argv[4]) if len(sys.argv) > 4 else 148 * 128
if rep.endswith(".csv"):          # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv > x_raw.csv
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe (tcgen05) %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem->tensor operand path %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]
nb = (n + 127) // 128
alg = {  # algorithmic work per launch
    "cov_build_kernel": ("bytes", 8.0 * 128 * 128 * nb * (nb + 1) / 2 + 24 * n),
    "chol_tiles_kernel": ("flop", n ** 3 / 3.0),
    "linv_tiles_kernel": ("flop", n ** 3 / 3.0),
    "trsv_forward_kernel": ("bytes", 4.0 * n * n),
    "trsv_backward_kernel": ("bytes", 4.0 * n * n),
    "var_tiles_kernel": ("flop", float(n) * n * q),
    "var_trsm_kernel": ("flop", float(n) * n * q),
    "predict_thread_kernel": ("pairs", float(n) * q),
}
lines = ["# ncu summary: %s" % rep.split("/")[-1], "",
         "One fit at n = %d + one mean+variance batch of %d queries (tools/prof_target.py), `ncu --set full "
         "--clock-control none`. Durations are cold-cache, serialised ncu replays: compare shares, not absolutes." % (n, q), "",
         "| kernel | " + " | ".join(w[1] for w in want) + " | algorithmic work | achieved |", "|---|" + "---|" * (len(want) + 2)]
n_chol = sum(1 for r in rows[2:] if "chol_tiles_kernel" in r[col["Kernel Name"]])
if n_chol > 1:        # INT8-assisted factorisation: the tile kernel only factorises panels, n^3/3 is the work of the whole sequence
    del alg["chol_tiles_kernel"]
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("gpr::", "")
    cells = []
    for m, _ in want:
        if m in col:
            v, u = r[col[m]], units[col[m]]
            try:
                v = "%.4g" % float(v)
            except ValueError:
                pass
            cells.append("%s %s" % (v, u) if u and u not in ("%", "register/thread") else v)
        else:
            cells.append("-")
    key = next((k for k in alg if k in short), None) if len(sys.argv) <= 5 else None
    work = ach = "-"
    if key:
        kind, amount = alg[key]
        t = float(r[col["gpu__time_duration.sum"]])
        tu = units[col["gpu__time_duration.sum"]]
        sec = t * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(tu, 1e-3)
        if kind == "bytes":
            work, ach = "%.3g GB" % (amount / 1e9), "%.0f GB/s" % (amount / sec / 1e9)
        elif kind == "flop":
            work, ach = "%.3g TFLOP" % (amount / 1e12), "%.1f TFLOP/s" % (amount / sec / 1e12)
        else:
            work, ach = "%.3g pair evals" % amount, "%.3g pairs/s" % (amount / sec)
    lines.append("| `%s` | %s | %s | %s |" % (short, " | ".join(cells), work, ach))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

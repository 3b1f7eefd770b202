#!/usr/bin/env python
"""bench.py — headline benchmark of the GP-regression hot path on B200 (BASELINE.json).

Workload (BASELINE config 3): synthetic cloud, seed 0, n = 16,384 (12,288 points on the unit
sphere, label 0; 4,096 on the r = 2 sphere, label +1), sigma2 = 0.1, ThinPlate(R = 4.2); queries are the
256^3 lattice on [-1.2, 1.2]^3, sharded over the ranks as contiguous index blocks (z-slabs).

A "step" is one pass of the hot path over one batch of queries per GPU: fused mean + cross-covariance
panel, then the variance product V = L^-1 K*^T (n^2 flop per query) in the library's default form for large batches: on
the INT8 tensor cores (tcgen05.mma kind::i8, TMEM accumulators, TMA feeds), FP64-equivalent by Ozaki slicing (6 slices of
base-254 digits per operand at this n, FP64 recombination; every call spot-checked against the FP64 tensor pipe).  The two FP64 forms
(forward substitution over L, product with L^-1 on the DMMA pipe) are timed on the same batches and reported beside it.  `value` is whole-job
query points / s (mean + variance) with the queries already resident in HBM; `e2e` is the same through
the host-pointer C-ABI call gpr_predict (pinned host buffers, H2D and D2H inside the timed region).
The fit (covariance build + Cholesky + alpha) is timed separately and reported as fit_ms on the same line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL broadcast of the model)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))

# stdout carries exactly ONE line, the JSON result: everything else that writes to file descriptor 1 (NCCL prints
# "NCCL version ..." there at communicator creation) is sent to stderr, and the result goes to the saved descriptor.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
sys.stdout.flush()
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()

sys.path.insert(0, ROOT)

N_TRAIN = 16384
GRID_RES = 256
BATCHES_PER_STEP = 4           # variance batches (148*128 queries each) per step and GPU
METRIC = "query_pts_per_sec_mean_var"
UNIT = "points/s"


def fit_roofline(extras, dmma_peak, dgemm):
    """The factorisation (n^3/3 flop).  All-FP64 tile kernel: against the DMMA issue peak.  INT8-assisted (the default for
    n >= 8192): ~83 % of the flops run as exact int8 slice products on tcgen05 (7 base-254 digits), the panels on the FP64 pipe
    — an FP64-equivalent rate, reported against the same FP64 peaks (it can exceed them; that is the point) and never clamped."""
    i8 = extras.get("fit_int8_slices", 0)
    r = {"kernel": ("launch_cholesky_int8: ozaki_var_kernel<%d,64,chunked,update> (INT8 tensor cores, left of each 16-tile panel) + "
                    "chol_tiles_kernel (FP64 tensor pipe, panels) + oz_slice_kernel; n^3/3 flop" % i8) if i8
         else "chol_tiles_kernel (tile-task Cholesky, n^3/3 flop)",
         "bound": "tensor", "achieved": extras["fit_chol_tflops"], "peak": dmma_peak,
         "unit": "TFLOP/s (FP64-equivalent)" if i8 else "TFLOP/s",
         "frac": extras["fit_chol_tflops"] / dmma_peak, "frac_vs_cublas": extras["fit_chol_tflops"] / dgemm,
         "ms": extras["fit_chol_ms"], "share_of_fit": extras["fit_chol_ms"] / extras["fit_ms"], "traffic": None,
         "int8_slices": i8}
    if i8:
        r["peak_note"] = "peak = measured FP64 DMMA issue rate; frac > 1 means the factorisation ran faster than any all-FP64 one can on this GPU"
    return r


def workload_config(extra=None):
    cfg = {"workload": "config3: synthetic sphere cloud n=16384, ThinPlate(R=4.2), sigma2=0.1, mean+variance over the "
                       "256^3 grid on [-1.2,1.2]^3 (z-slab shards)",
           "n_train": N_TRAIN, "grid": GRID_RES, "kernel": "thin_plate", "R": 4.2,
           "cache": "inputs larger than L2: each step streams the int8 slices of L^-1 (0.94 GB of 1.9 GB, lower triangle) and of a "
                    "2.5 GB cross-covariance panel (2.2 GB)"}
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_run(steps, warmup, sample_q=None):
    """The CPU arm, timed on this box's host cores.

    kind "reference": the reference's OWN evaluate(gp, query, f, v) (gp_regressor.hpp:282-324, compiled from
    /root/reference into oracle/_ref/libgpr_ref.so against the Eigen-API shim), driven the way the node drives it:
    one query per call from as many concurrent std::threads as there are cores (src/gp_node.cpp:1027-1038, :1074).
    Its model is installed from an OpenBLAS dpotrf factor, because the reference's unblocked single-threaded
    LDLT::compute takes tens of minutes at n = 16384 (that fit time is reported, not part of `value`).
    kind "port" (only when oracle/_ref is absent): the same mathematics with OpenBLAS dpotrf/dtrsm on all cores."""
    import oracle
    import gpr_b200
    W = gpr_b200.workloads
    P, y, s2 = W.synthetic_cloud(N_TRAIN, seed=0)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    model = oracle.blas_fit(P, y, s2, "thin_plate", W.SYNTH_R, 0.0)
    fit_s = time.perf_counter() - t0
    use_ref = oracle.have_reference()
    if sample_q is None:
        sample_q = 8 * cores if use_ref else 1024
    ref = None
    if use_ref:
        ref = oracle.Reference("thin_plate", W.SYNTH_R, 0.0)
        ref.adopt(P[:, 0], P[:, 1], P[:, 2], y, s2, model["alpha"], model["L"], 4.0)
    rng = np.random.default_rng(0)
    times, port_times = [], []
    for s in range(warmup + steps):
        z = int(rng.integers(0, GRID_RES))
        Q = W.grid_slab(GRID_RES, z, z + 1)
        Qs = Q[:sample_q]
        t0 = time.perf_counter()
        if use_ref:
            f, v = ref.evaluate_mt(Qs[:, 0], Qs[:, 1], Qs[:, 2], var=True, threads=cores, per_call=1)
        else:
            f, v = oracle.blas_predict(model, Qs, var=True)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        assert np.isfinite(f).all() and float(v.min()) > 0.0
    # the BLAS port on a 1024-query sample, for context next to the reference's own code
    t0 = time.perf_counter()
    oracle.blas_predict(model, W.grid_slab(GRID_RES, 7, 8)[:1024], var=True)
    port_qps = 1024 / (time.perf_counter() - t0)
    dt = sum(times)
    kind = "reference" if use_ref else "port"
    what = ("the reference's own evaluate(f, v), 1 query per call from %d concurrent threads" % cores) if use_ref else \
           "OpenBLAS port (dtrsm)"
    return {"last_queries": Qs, "last_f": f, "last_v": v, "alpha": model["alpha"],
            "value": sample_q * len(times) / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d steps x %d queries of the same grid (mean+variance), %s; model from one OpenBLAS dpotrf fit of "
                      "n=%d (%.1f s, not in value)" % (len(times), sample_q, what, N_TRAIN, fit_s),
            "fit_s": fit_s, "ms_per_step": 1e3 * dt / len(times), "sample_q": sample_q,
            "port_blas_dtrsm_points_per_s": port_qps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fit-reps", type=int, default=3)
    ap.add_argument("--no-full-grid", action="store_true", help="skip the strong-scaling pass over the whole 256^3 lattice")
    ap.add_argument("--no-fanout", action="store_true", help="skip the unchanged-caller thread fan-out measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        res = cpu_reference_run(args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(), "sample_queries_per_step": res["sample_q"],
                "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
                "fit_ms": 1e3 * res["fit_s"], "port_blas_dtrsm_points_per_s": res["port_blas_dtrsm_points_per_s"],
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import gpr_b200 as g
    from gaussian_object_modelling_b200 import distributed as D
    W = g.workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    ctx = g.Context(devices=[local_rank])
    reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
    extras = {}

    # ---- fit on rank 0 (the factorisation stays on one GPU), timed with the library's CUDA events ----
    model = None
    if rank == 0:
        P, y, s2 = W.synthetic_cloud(N_TRAIN, seed=0)
        fits = []
        for rep in range(1 + args.fit_reps):
            if model is not None:
                model.close()
            t0 = time.perf_counter()
            model = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
            wall = 1e3 * (time.perf_counter() - t0)
            t = ctx.timings()
            if rep == 0:
                # the very first fit of the process: 2.1 GB cudaMalloc, lazy module load, no buffer cache
                extras["fit_wall_ms_first_call"] = wall
            else:
                fits.append((t["fit_total_ms"], t["cov_ms"], t["chol_ms"], t["solve_ms"], wall))
                extras["fit_int8_slices"] = int(t["fit_int8_slices"])
        best = min(fits)
        extras.update(fit_ms=best[0], fit_cov_ms=best[1], fit_chol_ms=best[2], fit_solve_ms=best[3],
                      fit_wall_ms_e2e=min(f[4] for f in fits),      # host wall of gpr_fit (H2D of the cloud, allocation, D2H of alpha)
                      fit_chol_tflops=N_TRAIN ** 3 / 3 / (best[2] * 1e-3) / 1e12)
        # time to first variance: a fresh fit followed at once by one variance batch (148*128 queries), host wall — for the
        # default form (INT8 tensor cores: + one-time L^-1 and slicing) and for the forward substitution (needs nothing but L)
        sms0 = torch.cuda.get_device_properties(dev).multi_processor_count
        q1 = 128 * sms0
        Q1 = torch.from_numpy(np.ascontiguousarray(W.grid_slab(GRID_RES, 128, 129)[:q1].T)).to(dev)
        o1 = torch.empty(2 * q1, dtype=torch.float64, device=dev)
        ttfv = {}
        for mode in ("trsm", "default"):
            if mode == "trsm":
                os.environ["GPR_VAR_MODE"] = "trsm"
            else:
                os.environ.pop("GPR_VAR_MODE", None)
            model.close()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            model = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
            t_fit = 1e3 * (time.perf_counter() - t0)
            reg.evaluate_device(model, Q1[0].data_ptr(), Q1[1].data_ptr(), Q1[2].data_ptr(), q1, o1.data_ptr(), o1[q1:].data_ptr(), None)
            torch.cuda.synchronize(dev)
            tot = 1e3 * (time.perf_counter() - t0)
            tt = ctx.timings()
            ttfv[mode] = {"total_ms": tot, "fit_wall_ms": t_fit, "first_batch_queries": q1, "first_batch_wall_ms": tot - t_fit,
                          "of_which_linv_ms": tt["linv_ms"] if mode == "default" else 0.0, "variance_ms": tt["predict_var_ms"],
                          # the device-side part (CUDA events): what remains once the one-time multi-GB allocations of a fresh
                          # process (panel, L^-1, int8 slices: their host cost varies from box to box) are taken out
                          "device_ms": tt["fit_total_ms"] + (tt["linv_ms"] if mode == "default" else 0.0) + tt["predict_mean_ms"] + tt["predict_var_ms"]}
            if mode == "trsm":
                assert model.state().linv is None                # nothing built L^-1
        os.environ.pop("GPR_VAR_MODE", None)
        extras["time_to_first_variance_ms"] = ttfv["default"]["total_ms"]
        extras["time_until_variance_path_ready_ms"] = {"forward_substitution": ttfv["trsm"]["fit_wall_ms"],
                                                       "round1_flow_fit_plus_inverse": ttfv["default"]["fit_wall_ms"] + ttfv["default"]["of_which_linv_ms"]}
        extras["time_to_first_variance_split"] = ttfv
        del Q1, o1
    bcast_bytes = 0
    if world > 1:
        warm = torch.zeros(1 << 20, dtype=torch.float64, device=dev)
        dist.broadcast(warm, src=0)                        # communicator set-up is not part of the exchange step
        barrier()
        t0 = time.perf_counter()
        # the exchange step: {x|y|z, alpha, L, Dinv} — what the fit leaves behind; no L^-1 is built before it
        model, bcast_bytes = D.broadcast_model(reg, model, N_TRAIN, W.SYNTH_R, 2, rank, dev, src=0)
        barrier()
        bms = 1e3 * (time.perf_counter() - t0)
        extras.update(broadcast_ms=bms, broadcast_bytes=bcast_bytes, broadcast_GBps=bcast_bytes / bms / 1e6)
        # the same exchange fused into the producer: the replicas receive every tile of L and Dinv from inside the Cholesky
        # kernel (peer stores over NVLink); what remains exposed after the fit is the 32n-byte {x|y|z, alpha} broadcast
        model.close()
        model = None
        model, pinfo = D.fit_and_publish(reg, (lambda: reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)) if rank == 0 else None,
                                         N_TRAIN, W.SYNTH_R, rank, dev, src=0)
        if rank == 0:
            t = ctx.timings()
            pinfo.update(fit_ms=t["fit_total_ms"], chol_ms=t["chol_ms"], chol_ms_without_peers=extras["fit_chol_ms"])
            extras["fit_publish"] = pinfo
            extras["broadcast_exposed_ms"] = pinfo["exposed_ms"]

    # ---- this rank's block of the 256^3 query grid --------------------------------------------------
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    batch = 128 * sms
    step_q = BATCHES_PER_STEP * batch
    total_q = GRID_RES ** 3
    a, b = D.shard_range(total_q, rank, world)
    need = (args.warmup + args.steps) * step_q
    z0 = a // (GRID_RES * GRID_RES)
    nz = -(-need // (GRID_RES * GRID_RES)) + 1
    Qh = W.grid_slab(GRID_RES, z0, min(GRID_RES, z0 + nz))
    while len(Qh) < need:                                   # the shard is smaller than the run: wrap around
        Qh = np.vstack([Qh, Qh])
    Qh = np.ascontiguousarray(Qh[:need].T)                  # 3 x need, SoA like gp_regression::Data
    Qd = torch.from_numpy(Qh).to(dev)
    f_d = torch.empty(step_q, dtype=torch.float64, device=dev)
    v_d = torch.empty(step_q, dtype=torch.float64, device=dev)

    def step_device(s):
        o = s * step_q
        reg.evaluate_device(model, Qd[0, o:].data_ptr(), Qd[1, o:].data_ptr(), Qd[2, o:].data_ptr(), step_q,
                            f_d.data_ptr(), v_d.data_ptr(), None)
        t = ctx.timings()
        oz_ms[0] += t["ozaki_ms"]
        if t["ozaki_slices"] > 0:
            oz_used[0] = int(t["ozaki_slices"])
            oz_frac[0] = t["ozaki_issued_fraction"]
        return t["predict_var_ms"], t["predict_mean_ms"]

    oz_ms = [0.0]
    oz_used = [0]
    oz_frac = [1.0]
    sampler = ClockSampler(local_rank)
    for s in range(args.warmup):
        step_device(s)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    var_ms = mean_ms = 0.0
    oz_ms[0] = 0.0
    for s in range(args.warmup, args.warmup + args.steps):
        vm, mm = step_device(s)
        var_ms += vm
        mean_ms += mm
    e1.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = e0.elapsed_time(e1)
    oz_kernel_ms = oz_ms[0]
    if world > 1:
        tmax = torch.tensor([elapsed_ms, var_ms, mean_ms, oz_kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms, var_ms, mean_ms, oz_kernel_ms = (float(x) for x in tmax.tolist())
    value = world * step_q * args.steps / (elapsed_ms * 1e-3)
    assert bool(torch.isfinite(f_d).all()) and float(v_d.min()) > 0.0

    # ---- e2e: the host-pointer C-ABI call, pinned host buffers, copies inside the timed region -------
    Qp = torch.from_numpy(Qh).pin_memory()
    fo = torch.empty(step_q, dtype=torch.float64).pin_memory()
    vo = torch.empty(step_q, dtype=torch.float64).pin_memory()
    import ctypes as C
    dp = C.POINTER(C.c_double)

    def step_host(s):
        o = s * step_q
        ptr = lambda t, off=0: C.cast(t.data_ptr() + 8 * off, dp)
        rc = g.lib().gpr_predict(ctx._h, model._h, ptr(Qp[0], o), ptr(Qp[1], o), ptr(Qp[2], o), step_q, ptr(fo), ptr(vo), None, None, None)
        if rc:
            raise RuntimeError(g.lib().gpr_last_error().decode())

    for s in range(min(2, args.warmup)):
        step_host(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.warmup, args.warmup + args.steps):
        step_host(s)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * step_q * args.steps / e2e_s
    assert np.array_equal(fo.numpy(), f_d.cpu().numpy()) and np.array_equal(vo.numpy(), v_d.cpu().numpy())

    # ---- strong scaling: the WHOLE 256^3 lattice, fit -> last variance, sharded over the ranks (z-slabs) ----------
    # fit (rank 0) + broadcast + every rank's shard of the 16.8 M queries; wall clock of the slowest rank.
    full_grid = None
    if not args.no_full_grid:
        a_q, b_q = D.shard_range(total_q, rank, world)
        za, zb = a_q // (GRID_RES * GRID_RES), -(-b_q // (GRID_RES * GRID_RES))
        Qfull = torch.from_numpy(np.ascontiguousarray(W.grid_slab(GRID_RES, za, zb).T)).to(dev)
        off = a_q - za * GRID_RES * GRID_RES
        cnt = b_q - a_q
        fo_g = torch.empty(cnt, dtype=torch.float64, device=dev)
        vo_g = torch.empty(cnt, dtype=torch.float64, device=dev)
        barrier()
        t0 = time.perf_counter()
        fg_fit_ms = fg_bc_ms = 0.0
        if model is not None:
            model.close()
            model = None
        if world == 1:
            model = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
            fg_fit_ms = 1e3 * (time.perf_counter() - t0)
        else:
            model, pi2 = D.fit_and_publish(reg, (lambda: reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)) if rank == 0 else None,
                                           N_TRAIN, W.SYNTH_R, rank, dev, src=0)
            if rank == 0:
                fg_fit_ms, fg_bc_ms = pi2["fit_wall_ms"], pi2["exposed_ms"]
        reg.evaluate_device(model, Qfull[0, off:].data_ptr(), Qfull[1, off:].data_ptr(), Qfull[2, off:].data_ptr(), cnt,
                            fo_g.data_ptr(), vo_g.data_ptr(), None)
        torch.cuda.synchronize(dev)
        fg_s = time.perf_counter() - t0
        shell = int((fo_g.abs() <= 0.01).sum().item())
        vmin = float(vo_g.min().item())
        if world > 1:
            tt = torch.tensor([fg_s, fg_fit_ms, fg_bc_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            fg_s, fg_fit_ms, fg_bc_ms = (float(x) for x in tt.tolist())
            cc = torch.tensor([float(shell), -vmin], dtype=torch.float64, device=dev)
            dist.all_reduce(cc[:1], op=dist.ReduceOp.SUM)
            dist.all_reduce(cc[1:], op=dist.ReduceOp.MAX)
            shell, vmin = int(cc[0].item()), -float(cc[1].item())
        assert vmin > 0.0
        full_grid = {"full_grid_s": fg_s, "queries": total_q, "points_per_s": total_q / fg_s, "fit_wall_ms": fg_fit_ms,
                     "broadcast_ms": fg_bc_ms, "points_in_shell_abs_f_le_0.01": shell, "scaling": "strong",
                     "what": ("fit on rank 0" if world == 1 else
                              "replica set-up (IPC handles) -> fit on rank 0 publishing L and Dinv into the replicas from inside the "
                              "factorisation -> {x|y|z, alpha} broadcast" if D.publish_pays(N_TRAIN, world - 1) else
                              "fit on rank 0 -> NCCL broadcast of {x|y|z, alpha, L, Dinv} (publishing %d copies from inside the "
                              "factorisation would not hide in it: distributed.publish_pays)" % (world - 1))
                             + " -> mean+variance of ALL 256^3 lattice points (z-slab shards); host wall of the slowest rank"}
        del Qfull, fo_g, vo_g

    if rank == 0:
        # the other evaluate overloads on this GPU (device-resident, CUDA-event timed by the library), for context
        qn = min(need, 1 << 20)
        f_x = torch.empty(qn, dtype=torch.float64, device=dev)
        g_x = torch.empty(3 * qn, dtype=torch.float64, device=dev)
        others = {}
        for name, gp, vp, nq in (("mean_only", None, None, qn), ("mean_grad", g_x.data_ptr(), None, qn),
                                 ("mean_var_grad", g_x.data_ptr(), v_d.data_ptr(), step_q)):
            best = None
            for _ in range(3):
                reg.evaluate_device(model, Qd[0].data_ptr(), Qd[1].data_ptr(), Qd[2].data_ptr(), nq, f_x.data_ptr(), vp, gp)
                t = ctx.timings()
                ms = t["predict_mean_ms"] + t["predict_var_ms"]
                best = ms if best is None or ms < best else best
            others[name + "_pts_per_s_1gpu"] = nq / (best * 1e-3)
        others["mean_only_pair_evals_per_s"] = others["mean_only_pts_per_s_1gpu"] * N_TRAIN
        extras["other_overloads"] = others
        launches_var = BATCHES_PER_STEP * args.steps
        flops_per_launch = float(N_TRAIN) ** 2 * batch                # n^2 * q per variance batch, FP64 count
        fp64_equiv = flops_per_launch / (var_ms / launches_var * 1e-3) / 1e12
        # the INT8 kernel's own work: 2 ops x (slice pairs t + u < S) x (lower triangle by 128-row tiles) x queries
        oz_slices = oz_used[0] if oz_used[0] else 6
        oz_pairs = oz_slices * (oz_slices + 1) // 2
        int8_ops_dense = 2.0 * oz_pairs * (N_TRAIN * (N_TRAIN + 128) / 2.0) * batch
        # digit slices of L^-1 that are all zero in a (128-row, 64-k) block are skipped: only the issued MMAs count as work done
        int8_ops_per_launch = int8_ops_dense * oz_frac[0]
        default_is_int8 = oz_kernel_ms > 0.0
        achieved = int8_ops_per_launch / (oz_kernel_ms / launches_var * 1e-3) / 1e12 if default_is_int8 else fp64_equiv
        # the two FP64 forms on the same batches (forced through the environment, read per call)
        forms = {}
        for mode in ("product", "trsm"):
            os.environ["GPR_VAR_MODE"] = mode
            if mode == "product":
                reg.prepare_variance(model)
            pv = [step_device(s_)[0] for s_ in range(3)]
            forms[mode] = {"var_ms_per_step": min(pv), "fp64_tflops": flops_per_launch * BATCHES_PER_STEP / (min(pv) * 1e-3) / 1e12,
                           "pts_per_s": step_q / ((min(pv) + mean_ms / args.steps) * 1e-3)}
        os.environ.pop("GPR_VAR_MODE", None)
        forms["product"]["kernel"] = "var_tiles_kernel (product with the explicit inverse factor, DMMA)"
        forms["trsm"]["kernel"] = "var_trsm_kernel (blocked forward substitution over L in the K* panel, DMMA; needs no inverse)"
        forms["int8_default" if default_is_int8 else "default"] = {
            "var_ms_per_step": var_ms / args.steps, "fp64_equivalent_tflops": fp64_equiv, "pts_per_s": value / world,
            "kernel": "ozaki_var_kernel<%d> (tcgen05.mma kind::i8 + TMEM + TMA)" % oz_slices, "kernel_ms_per_step": oz_kernel_ms / args.steps,
            "other_ms_per_step": "panel slicing (oz_slice_kernel), finalize and the per-call FP64 spot check: %.2f" % ((var_ms - oz_kernel_ms) / args.steps)}
        extras["variance_forms"] = forms
        # The denominators, reported separately and never clamped.  INT8: MEASURED_PEAKS.json holds no int8 entry; the int8
        # tensor rate is twice the bf16 rate on this part (4.5 vs 2.25 PFLOP/s nominal), so the peak used is 2 x the measured
        # sustained bf16 figure (the kernel is timed inside a long step), with the nominal 4500 beside it.  FP64: (1) the raw
        # DMMA issue rate of this device (gpr_selftest_peak), best of several readings; (2) cuBLAS DGEMM 8192^3.
        readings = []
        for attempt in range(6):
            readings.append(g.selftest_peak(0, 4))
            time.sleep(0.4)
        dmma_peak = max(readings)
        mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp_path):
            mp = json.load(open(mp_path))
            bf16_sus, bf16_burst, peak_kind = mp.get("bf16_tflops_sustained", 1350.8), mp.get("bf16_tflops", 1595.1), "of measured"
        else:
            bf16_sus, bf16_burst, peak_kind = 1400.0, 1590.0, "of fallback"
        int8_peak = 2.0 * bf16_sus
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ozaki_traffic.json" if default_is_int8 else "var_trsm_traffic.json")
        if os.path.exists(tpath):          # dram__bytes_read+write of one launch, from the committed ncu --set full capture
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        a64 = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        for _ in range(2):
            a64 @ a64
        torch.cuda.synchronize(dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(); a64 @ a64; a64 @ a64; c1.record(); torch.cuda.synchronize(dev)
        dgemm = 2 * 2 * 8192 ** 3 / (c0.elapsed_time(c1) * 1e-3) / 1e12
        del a64
        peak_src = ("INT8 tensor cores: 2 x bf16_tflops_sustained of MEASURED_PEAKS.json (%s; %.1f TF/s bf16 sustained, %.1f burst; no int8 "
                    "entry there; nominal int8 dense 4500 TOP/s).  FP64 tensor pipe for the fp64-equivalent figures: raw DMMA.8x8x4 issue "
                    "rate measured in this run (gpr_selftest_peak), best of %d readings %s; cuBLAS DGEMM 8192^3 in this run: %.1f TF/s"
                    % (peak_kind, bf16_sus, bf16_burst, len(readings), ["%.1f" % r for r in readings], dgemm))
        if default_is_int8:
            roof = {"kernel": "ozaki_var_kernel<%d> (variance product X K*^T on the INT8 tensor cores: tcgen05.mma kind::i8, TMEM int32 "
                              "accumulators, TMA feeds; FP64 recombination + column norms in the epilogue)" % oz_slices,
                    "bound": "tensor", "achieved": achieved, "peak": int8_peak, "unit": "TOP/s (int8 multiply-adds x 2)",
                    "frac": achieved / int8_peak, "peak_nominal": 4500.0, "frac_of_nominal": achieved / 4500.0,
                    "algorithmic_int8_ops_per_launch": int8_ops_per_launch, "issued_fraction_of_dense_slice_pairs": oz_frac[0],
                    "dense_int8_ops_per_launch": int8_ops_dense, "slices": oz_slices, "slice_pairs": oz_pairs,
                    "digit_base": 254 if oz_slices <= 6 else 128,
                    "fp64_equivalent_tflops": fp64_equiv, "fp64_equivalent_vs_dmma_peak": fp64_equiv / dmma_peak,
                    "fp64_equivalent_vs_cublas_dgemm": fp64_equiv / dgemm}
        else:
            roof = {"kernel": "var_trsm_kernel", "bound": "tensor", "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
                    "frac": achieved / dmma_peak, "frac_vs_cublas": achieved / dgemm}
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "algorithmic_flop_per_launch_fp64": flops_per_launch,
                     "algorithmic_operand_bytes": (oz_slices if default_is_int8 else 8.0) * (N_TRAIN * (N_TRAIN + 128) / 2 + N_TRAIN * batch),
                     "peak_source": peak_src, "cublas_dgemm_tflops": dgemm, "dmma_probe_tflops": dmma_peak,
                     "share_of_step": (oz_kernel_ms if default_is_int8 else var_ms) / elapsed_ms,
                     "mean_panel_kernel_ms_per_step": mean_ms / args.steps,
                     "fp64_forms": {"var_trsm_kernel": {"achieved": forms["trsm"]["fp64_tflops"], "peak": dmma_peak, "unit": "TFLOP/s",
                                                        "frac": forms["trsm"]["fp64_tflops"] / dmma_peak},
                                    "var_tiles_kernel": {"achieved": forms["product"]["fp64_tflops"], "peak": dmma_peak, "unit": "TFLOP/s",
                                                         "frac": forms["product"]["fp64_tflops"] / dmma_peak}}})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 (int8-sliced: exact int8 x int8 -> int32 tensor-core products, FP64 recombination)" if default_is_int8 else "f64",
                "data": "synthetic", "config": workload_config(),
                "queries_per_step_per_gpu": step_q, "parallelism": "query-sharded x%d" % world,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24 * step_q, "d2h_bytes_per_step": 16 * step_q},
                # per variance batch: predict_thread_kernel (mean + K* panel), predict_reduce_kernel, oz_slice_kernel, ozaki_var_kernel,
                # var_finalize_kernel; per call: the spot check (var_tiles_kernel + var_finalize_kernel)
                "gpu_launches": (5 * BATCHES_PER_STEP + 2) * args.steps if default_is_int8 else 4 * BATCHES_PER_STEP * args.steps,
                "clocks": clocks,
                "roofline": roof,
                "roofline_fit": fit_roofline(extras, dmma_peak, dgemm)}
        line.update(extras)
        if full_grid:
            line["full_grid"] = full_grid
            line["full_grid_s"] = full_grid["full_grid_s"]
        if world == 1 and not args.no_fanout:
            # the unchanged caller: 29 slabs x 841 std::threads x one evaluate(q = 1) (src/gp_node.cpp:1025-1038) through the
            # drop-in headers, and the same source against the reference's own header on this box's host cores
            try:
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fanout_bench.py"), "--cases", "mugD:node"],
                                     capture_output=True, text=True, timeout=900)
                fj = json.loads(out.stdout.strip().splitlines()[-1])["cases"][0]
                line["fanout_calls_per_s"] = fj["ours"]["calls_per_s"]
                line["fanout"] = {"workload": "mugD (n = 277), ThinPlate(2.0) — the node's own setting; 24389 single-query "
                                              "evaluate(f, v) calls from 841 concurrent std::threads per slab",
                                  "ours": fj["ours"], "reference_cpu": fj.get("reference"), "parity": fj.get("parity"),
                                  "speedup_vs_reference": fj.get("speedup_vs_reference")}
            except Exception as e:                      # the headline line must not depend on g++ being present
                line["fanout"] = {"error": repr(e)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            res = cpu_reference_run(2, 1)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "fit_s",
                                                        "port_blas_dtrsm_points_per_s")}
            # parity at the full bench size, on the CPU arm's last sample: our alpha against the OpenBLAS
            # Cholesky solve, our mean / variance against what the reference's evaluate() returned
            # (tests/test_gpu_parity.py::test_headline_parity_config3 / _config5 hold the asserted versions)
            Qs = res["last_queries"]
            fg, vg = reg.evaluate(model, Qs[:, 0], Qs[:, 1], Qs[:, 2], var=True)
            rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
            big = np.abs(res["last_f"]) > 1e-9
            line["parity_full_size"] = {"against": res["kind"], "queries": int(len(Qs)),
                                        "alpha_rel_inf_vs_dpotrs": rel(model.alpha, res["alpha"]),
                                        "mean_rel_inf": rel(fg, res["last_f"]), "var_rel_inf": rel(vg, res["last_v"]),
                                        "sign_mismatches": int((np.sign(fg) != np.sign(res["last_f"]))[big].sum()),
                                        "tolerance": {"alpha": "max(1e-9, 50*cond*eps)", "mean": 1e-9, "var": 1e-7}}
        emit(line)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""BASELINE config 3 end to end under torchrun: fit n = 16,384 on rank 0, broadcast {x|y|z, alpha, L^-1} once, then EVERY
rank evaluates its z-slab block of the complete 256^3 lattice (mean + variance).  Rank 0 prints one JSON line.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/full_grid.py [grid=256] [n=16384] [chunks_per_rank]
With chunks_per_rank > 0 every rank evaluates only that many batches of its block (config 5: n = 65,536, grid 512 — the
complete sweep is extrapolated from the slab)."""
import json
import os
import sys
import time

import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import gpr_b200 as g
from gaussian_object_modelling_b200 import distributed as D

W = g.workloads
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_train = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
max_chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # > 0: only this many chunks per rank (a slab; config 5)
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    dist.broadcast(torch.zeros(1 << 20, dtype=torch.float64, device=dev), src=0)
ctx = g.Context(devices=[local])
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
model, fit_ms, linv_ms = None, 0.0, 0.0
t_all = time.perf_counter()
if rank == 0:
    P, y, s2 = W.synthetic_cloud(n_train, seed=0)
    model = reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)
    fit_ms = ctx.timings()["fit_total_ms"]
    reg.prepare_variance(model)
    linv_ms = ctx.timings()["linv_ms"]
t0 = time.perf_counter()
if world > 1:
    model, nbytes = D.broadcast_model(reg, model, n_train, W.SYNTH_R, True, rank, dev, src=0)
    dist.barrier(device_ids=[local])
bcast_ms = 1e3 * (time.perf_counter() - t0)
total = res ** 3
a, b = D.shard_range(total, rank, world)
chunk = 148 * 128 * (8 if n_train <= 16384 else 1)
if max_chunks > 0:
    b = min(b, a + max_chunks * chunk)
done = b - a
fmin, vmin, vmax, shell, dev_ms = np.inf, np.inf, -np.inf, 0, 0.0
torch.cuda.synchronize(dev)
t0 = time.perf_counter()
f_d = torch.empty(chunk, dtype=torch.float64, device=dev)
v_d = torch.empty(chunk, dtype=torch.float64, device=dev)
lin = torch.linspace(-W.SYNTH_GRID_HALF, W.SYNTH_GRID_HALF, res, dtype=torch.float64, device=dev)
for s in range(a, b, chunk):
    e = min(b, s + chunk)
    idx = torch.arange(s, e, device=dev)
    qz, qy, qx = lin[idx // (res * res)], lin[(idx // res) % res], lin[idx % res]       # z-major slabs, like workloads.grid_slab
    q = e - s
    reg.evaluate_device(model, qx.contiguous().data_ptr(), qy.contiguous().data_ptr(), qz.contiguous().data_ptr(), q, f_d.data_ptr(), v_d.data_ptr(), None)
    t = ctx.timings()
    dev_ms += t["predict_mean_ms"] + t["predict_var_ms"]
    fmin = min(fmin, float(f_d[:q].abs().min())); vmin = min(vmin, float(v_d[:q].min())); vmax = max(vmax, float(v_d[:q].max()))
    shell += int((f_d[:q].abs() <= 0.01).sum())
torch.cuda.synchronize(dev)
wall = time.perf_counter() - t0
stats = torch.tensor([wall, dev_ms, -vmin, vmax, float(shell), float(done)], dtype=torch.float64, device=dev)
if world > 1:
    mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    wall, dev_ms, vmin, vmax, shell, done = float(mx[0]), float(mx[1]), -float(mx[2]), float(mx[3]), int(sm[4]), int(sm[5])
if rank == 0:
    print(json.dumps({"workload": "n=%d fit + %s of the %d^3 = %d lattice, mean+variance, z-slab shards" % (
                          n_train, "ALL" if max_chunks == 0 else "%d batches per rank" % max_chunks, res, total),
                      "n_gpus": world, "fit_ms": fit_ms, "linv_ms": linv_ms, "broadcast_ms": bcast_ms,
                      "predict_wall_s_max_over_ranks": wall, "predict_device_s_max_over_ranks": dev_ms / 1e3,
                      "points_evaluated": int(done), "points_per_s_whole_job": done / wall,
                      "extrapolated_full_lattice_s": total / (done / wall), "fit_to_last_variance_s": time.perf_counter() - t_all,
                      "var_min": vmin, "var_max": vmax, "points_with_abs_f_le_0.01": shell}))
if world > 1:
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()

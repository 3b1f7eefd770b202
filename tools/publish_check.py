"""Under torchrun (2+ ranks): the exchange step fused into the fit.  The other ranks create their replicas first and export
CUDA IPC handles; rank 0's Cholesky kernel stores every finished tile of L and Dinv into them while it factorises
(distributed.fit_and_publish); then every rank evaluates its shard of a query set by forward substitution over ITS copy of
the factor, and rank 0 compares the gathered mean / variance bit for bit with its own evaluation of all queries, and with the
legacy path (factor broadcast with NCCL after the fit).  Also: an indefinite matrix must fall back to the broadcast of L^-1
and the tail block.  Prints PUBLISH_OK on success."""
import os
import sys
import numpy as np
os.environ["GPR_FIT_PUBLISH"] = "1"      # this tool checks the mechanism; distributed.publish_pays would decline at this size
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import gpr_b200 as g
from gaussian_object_modelling_b200 import distributed as D

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W = g.workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000             # ragged: padded to 3072
P, y, s2 = W.synthetic_cloud(n, seed=5)
Q = W.grid_slab(64, 20, 24)                                      # 16384 queries: the forward-substitution path on every rank
ctx = g.Context(devices=[local])
reg = g.GPRegressor("thin_plate", W.SYNTH_R, ctx=ctx)
model, info = D.fit_and_publish(reg, (lambda: reg.create(P[:, 0], P[:, 1], P[:, 2], y, s2)) if rank == 0 else None,
                                n, W.SYNTH_R, rank, dev, src=0)
assert info["published"]
a, b = D.shard_range(len(Q), rank, world)
os.environ["GPR_VAR_MODE"] = "trsm"                              # whatever the shard size: the form that needs only the factor
f, v = reg.evaluate(model, Q[a:b, 0], Q[a:b, 1], Q[a:b, 2], var=True)
assert model.state().linv is None                                # no inverse factor anywhere
fs = [None] * world; vs = [None] * world
dist.all_gather_object(fs, f); dist.all_gather_object(vs, v)
# legacy path for comparison: broadcast of the factor after the fit
legacy = model if rank == 0 else None
legacy, nbytes = D.broadcast_model(reg, legacy, n, W.SYNTH_R, 2, rank, dev, src=0)
f2, v2 = reg.evaluate(legacy, Q[a:b, 0], Q[a:b, 1], Q[a:b, 2], var=True)
assert np.array_equal(f, f2) and np.array_equal(v, v2)
ok = True
if rank == 0:
    F, V = np.concatenate(fs), np.concatenate(vs)
    f0, v0 = reg.evaluate(model, Q[:, 0], Q[:, 1], Q[:, 2], var=True)
    same = np.array_equal(F, f0) and np.array_equal(V, v0)
    print("publish: n=%d world=%d fit %.2f ms exposed %.3f ms factor %.1f MB/peer identical=%s"
          % (n, world, info["fit_wall_ms"], info["exposed_ms"], info["factor_bytes_per_peer"] / 1e6, same))
    ok = same and v0.min() > 0
os.environ.pop("GPR_VAR_MODE", None)
# an indefinite matrix (the node's setting): nothing is published, the state is broadcast the old way
z = np.load(os.path.join(ROOT, "tests", "golden", "ref_mugD_thinplate_R2_node.npz"))
Pn, Qn = z["P"], np.vstack([z["Q"]] * 8)
reg2 = g.GPRegressor("thin_plate", 2.0, ctx=ctx)
m2, info2 = D.fit_and_publish(reg2, (lambda: reg2.create(Pn[:, 0], Pn[:, 1], Pn[:, 2], z["y"], z["s2"])) if rank == 0 else None,
                              len(Pn), float(z["R"]), rank, dev, src=0)
assert not info2["published"] and m2.n_tail == 15
a, b = D.shard_range(len(Qn), rank, world)
fn, vn = reg2.evaluate(m2, Qn[a:b, 0], Qn[a:b, 1], Qn[a:b, 2], var=True)
q = len(z["Q"])
ref_f, ref_v = np.concatenate([z["f"]] * 8)[a:b], np.concatenate([z["v"]] * 8)[a:b]
assert np.abs(fn - ref_f).max() <= 1e-9 * np.abs(z["f"]).max() and np.abs(vn - ref_v).max() <= 1e-7 * np.abs(z["v"]).max()
flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0 and flag.item() == 1.0:
    print("PUBLISH_OK")
dist.barrier(device_ids=[local])
dist.destroy_process_group()

"""Under torchrun (2+ ranks): rank 0 fits the node's indefinite configuration (mugD + external points, ThinPlate(2.0)),
the model incl. its tail block is broadcast, every rank evaluates its shard of the fixture's queries; rank 0 compares the
gathered mean / variance with the reference fixture.  Prints TAIL_REPLICA_OK on success."""
import os
import sys
import numpy as np
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
z = np.load(os.path.join(ROOT, "tests", "golden", "ref_mugD_thinplate_R2_node.npz"))
P, Q = z["P"], np.vstack([z["Q"]] * 40)                      # 10,080 queries: the tile-variance path on every rank
ctx = g.Context(devices=[local])
reg = g.GPRegressor("thin_plate", 2.0, ctx=ctx)
model = None
if rank == 0:
    model = reg.create(P[:, 0], P[:, 1], P[:, 2], z["y"], z["s2"])
    assert model.n_tail == 15
    reg.prepare_variance(model)
model, nbytes = D.broadcast_model(reg, model, len(P), float(z["R"]), True, rank, dev, src=0)
a, b = D.shard_range(len(Q), rank, world)
f, v = reg.evaluate(model, Q[a:b, 0], Q[a:b, 1], Q[a:b, 2], var=True)
fs = [None] * world; vs = [None] * world
dist.all_gather_object(fs, f); dist.all_gather_object(vs, v)
if rank == 0:
    F, V = np.concatenate(fs), np.concatenate(vs)
    q = len(z["Q"])
    ef = np.abs(F[:q] - z["f"]).max() / np.abs(z["f"]).max()
    ev = np.abs(V[:q] - z["v"]).max() / np.abs(z["v"]).max()
    same = np.array_equal(F[:q], F[-q:]) and np.array_equal(V[:q], V[-q:])   # first block on rank 0, last block on the last rank
    print("tail replica: n_tail=%d bytes=%d mean %.2e var %.2e shards identical=%s" % (model.n_tail, nbytes, ef, ev, same))
    assert ef <= 1e-9 and ev <= 1e-7 and same
    print("TAIL_REPLICA_OK")
dist.barrier(device_ids=[local])
dist.destroy_process_group()

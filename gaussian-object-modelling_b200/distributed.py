"""One-process-per-GPU plumbing for the query-sharded predict.

The path shards by query point and has exactly one exchange step: after the fit on rank 0, the state
that predict needs — X (3·N doubles), alpha (N) and, for variances, the inverse factor L^-1 (N² doubles)
— is broadcast once (NCCL over NVLink on the GPU box, gloo in the CPU tests).  There is no per-query
communication; every query is computed by exactly one rank with identical code, so sharded results are
bit-identical to single-GPU results.  torch.distributed is plumbing only: the buffers being broadcast
are the C-ABI library's own device allocations, wrapped without a copy.
"""
import numpy as np


def shard_range(q, rank, world):
    """Contiguous block of the query index range owned by `rank` (z-slabs for grids)."""
    return (q * rank) // world, (q * (rank + 1)) // world


class _DevicePtr:
    """Zero-copy view of a raw device allocation for torch.as_tensor (CUDA array interface v3)."""

    def __init__(self, ptr, n_doubles):
        self.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def as_tensor(ptr, n_doubles, device):
    import torch
    return torch.as_tensor(_DevicePtr(ptr, n_doubles), device=device)


def broadcast_model(reg, model, n, R, with_linv, rank, device, src=0):
    """Rank `src` passes its fitted `model`; the other ranks pass model=None and receive a replica.
    with_linv: False/0 mean only; True/1 the inverse factor L^-1 (built on `src` if it was not yet); 2 the Cholesky
    factor L + the inverses of its diagonal blocks (what a fit leaves behind: nothing to build before the exchange;
    the replicas then compute large-batch variances by forward substitution).  Returns (model, bytes_broadcast)."""
    import torch
    import torch.distributed as dist
    n_tail = model.n_tail if rank == src else 0
    meta = torch.tensor([float(n), float(R), float(n_tail)], dtype=torch.float64, device=device)
    dist.broadcast(meta, src=src)
    n, R, n_tail = int(meta[0].item()), float(meta[1].item()), int(meta[2].item())
    if rank != src:
        model = reg.create_replica(n, R, with_linv, n_tail)
    st = model.state(with_linv=with_linv)
    N = st.padded_n
    if st.ld != N:
        raise ValueError("model has spare capacity (ld %d != padded n %d): broadcast a freshly fitted model" % (st.ld, N))
    nbytes = 0
    with_linv = int(with_linv)
    parts = ((st.xyz, 3 * N), (st.alpha, N)) + (((st.linv, N * N),) if with_linv & 1 else ())
    if with_linv & 2:
        if not st.lfac:
            raise ValueError("model holds no Cholesky factor to broadcast (indefinite tail block?)")
        parts += ((st.lfac, N * N), (st.dinv, (N // 128) * 128 * 128))
    if n_tail:                                  # indefinite tail block: Z = A^-1 P (slabs) and S^-1
        parts += ((st.tail_z, (st.tail_pad // 32) * N * 32), (st.tail_sinv, st.tail_pad * st.tail_pad))
    for ptr, cnt in parts:
        t = as_tensor(ptr, cnt, device)
        # NCCL counts are 32-bit element counts in some paths: broadcast the factor in 1 GiB pieces
        step = 1 << 27
        for a in range(0, cnt, step):
            dist.broadcast(t[a:a + step], src=src)
        nbytes += 8 * cnt
    torch.cuda.synchronize(device)
    return model, nbytes


def publish_pays(n, peers):
    """Publishing pushes `peers` copies of the factor (8 n^2 bytes each) through the fitting GPU's NVLink egress while the
    factorisation runs; it pays while that stays well below the factorisation time, otherwise one NCCL broadcast after the
    fit is cheaper.  Measured at n = 16 384 with the INT8-assisted fit (26 ms): 1 peer +2.4 ms in the fit and 0.4 ms exposed
    against a 5.6 ms broadcast; 7 peers +21 ms against a 12.3 ms broadcast.  The transfer grows as peers n^2, the fit as n^3:
    publish while peers <= n / 5500 (n = 16 384: up to 2 peers; n = 65 536: 11).  GPR_FIT_PUBLISH=0|1 overrides."""
    import os
    env = os.environ.get("GPR_FIT_PUBLISH")
    if env is not None:
        return env != "0"
    return peers >= 1 and peers * 5500 <= n


def fit_and_publish(reg, fit, n, R, rank, device, src=0):
    """The exchange step fused into the fit (when publish_pays(n, peers); otherwise fit, then broadcast the factor).  Every rank but `src` creates its replica FIRST and exports CUDA IPC handles
    of its factor buffers; `src` registers them (gpr_ctx_set_fit_peers) and then runs `fit()` (a callable returning the
    fitted Model): its Cholesky kernel stores every finished tile of L and Dinv into all replicas over NVLink while it
    factorises, so when the fit returns only {x|y|z, alpha} (32 n bytes) are left to broadcast.  If the matrix turns out
    indefinite (trailing-block path) nothing was published and the state is broadcast the old way.
    Returns (model, info) with info = {published, fit_wall_ms (src), exposed_ms: src's fit end -> every rank ready}."""
    import time
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    if not publish_pays(n, world - 1):
        dist.barrier(device_ids=[device.index])
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        model = fit() if rank == src else None
        t_fit = time.perf_counter()
        tail = torch.tensor([float(model.n_tail) if rank == src else 0.0], dtype=torch.float64, device=device)
        dist.broadcast(tail, src=src)
        model, nbytes = broadcast_model(reg, model, n, R, 1 if tail.item() else 2, rank, device, src=src)
        dist.barrier(device_ids=[device.index])
        torch.cuda.synchronize(device)
        t_end = time.perf_counter()
        return model, {"published": False, "policy": "broadcast after the fit (publishing %d copies would not hide in the fit)" % (world - 1),
                       "fit_wall_ms": 1e3 * (t_fit - t0) if rank == src else None,
                       "exposed_ms": 1e3 * (t_end - t_fit) if rank == src else None, "small_state_bytes": 0,
                       "factor_bytes_per_peer": nbytes}
    replica, blob = None, None
    if rank != src:
        replica = reg.create_replica(n, R, 2)
        blob = replica.ipc_export()
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    if rank == src:
        reg.ctx.set_fit_peers([b for r, b in enumerate(blobs) if r != src], n)
    dist.barrier(device_ids=[device.index])
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    model, published, n_tail = replica, 0.0, 0.0
    if rank == src:
        model = fit()
        published, n_tail = float(reg.ctx.last_fit_published), float(model.n_tail)
    t_fit = time.perf_counter()
    meta = torch.tensor([published, n_tail], dtype=torch.float64, device=device)
    dist.broadcast(meta, src=src)
    published = bool(meta[0].item())
    nbytes = 0
    if published:
        st = model.state(with_linv=0)
        N = st.padded_n
        for ptr, cnt in ((st.xyz, 3 * N), (st.alpha, N)):
            dist.broadcast(as_tensor(ptr, cnt, device), src=src)
            nbytes += 8 * cnt
    else:
        if rank != src:
            model.close()
            model = None
        model, nbytes = broadcast_model(reg, model, n, R, 1, rank, device, src=src)
    dist.barrier(device_ids=[device.index])
    torch.cuda.synchronize(device)
    t_end = time.perf_counter()
    if rank == src:
        reg.ctx.clear_fit_peers()
    info = {"published": published, "fit_wall_ms": 1e3 * (t_fit - t0) if rank == src else None,
            "exposed_ms": 1e3 * (t_end - t_fit) if rank == src else None, "small_state_bytes": nbytes,
            "factor_bytes_per_peer": 8 * (((n + 127) // 128 * 128) ** 2 + ((n + 127) // 128) * 128 * 128)}
    return model, info


def broadcast_arrays(arrays, src=0):
    """gloo/CPU counterpart used by the tests: broadcast a list of numpy float64 arrays in place."""
    import torch
    import torch.distributed as dist
    for a in arrays:
        t = torch.from_numpy(a)
        dist.broadcast(t, src=src)
    return arrays


def gather_shards(local, q, rank, world):
    """All-gather variable-length query shards back into one array of length q (tests only)."""
    import torch
    import torch.distributed as dist
    sizes = [shard_range(q, r, world)[1] - shard_range(q, r, world)[0] for r in range(world)]
    m = max(sizes)
    buf = np.zeros(m)
    buf[:len(local)] = local
    outs = [torch.zeros(m, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(outs, torch.from_numpy(buf))
    return np.concatenate([o.numpy()[:s] for o, s in zip(outs, sizes)])

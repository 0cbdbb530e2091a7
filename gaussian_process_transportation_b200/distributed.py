"""Query sharding across GPUs: the fit runs on one rank, the immutable model state
(X, alpha, inverse factor, hyper-parameters) is broadcast once with NCCL over NVLink, queries are block-partitioned and
need no further exchange.  One process per GPU, `torch.distributed` is the plumbing."""
from __future__ import annotations

import numpy as np


def shard_bounds(M: int, world: int, rank: int):
    """Contiguous block partition of M queries: ranks [0, M % world) get one extra."""
    base, extra = divmod(int(M), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _DevBuf:
    """Zero-copy view of an engine-owned device buffer for torch (`__cuda_array_interface__`)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def broadcast_tensors(tensors, src=0, group=None):
    import torch.distributed as dist
    for t in tensors:
        dist.broadcast(t, src=src, group=group)


def broadcast_model(gp_or_engine, shape=None, src=0, group=None, with_variance=True):
    """Broadcast a fitted model from `src` to every rank.  On `src` pass the fitted GaussianProcess / Engine; on the
    other ranks pass an un-fitted one: its ENGINE receives the state (X, alpha, inverse factor, hyper-parameters, ordering) and serves
    `Engine.query / query_dev / query_grid` for that rank's shard.  The host-side object of a receiving rank stays un-fitted (no kernel
    object, no training arrays travel): receivers query through the returned engine, not through `predict` / `derivative`.
    Returns the engine."""
    import torch
    import torch.distributed as dist
    eng = getattr(gp_or_engine, "_engine", gp_or_engine)
    rank = dist.get_rank(group)
    dev = torch.device("cuda", eng.device)
    meta = torch.zeros(5, dtype=torch.int64, device=dev)
    if rank == src:
        if with_variance:
            eng.prepare_variance()
        meta[:] = torch.tensor([eng.N, eng.d, eng.p, int(with_variance), int(eng.spatial)], dtype=torch.int64)
    dist.broadcast(meta, src=src, group=group)
    N, d, p, wv, sp = (int(v) for v in meta.tolist())
    if rank != src:
        eng.state_alloc(N, d, p, bool(wv))
        eng.spatial = bool(sp)                    # the handle itself learns the ordering from the state header (gptb_state_commit)
    bufs = []
    for which in (0, 1, 2) + ((3,) if wv else ()):
        ptr, nbytes = eng.state_buffer(which)
        bufs.append(torch.as_tensor(_DevBuf(ptr, nbytes), device=dev))
    broadcast_tensors(bufs, src=src, group=group)
    torch.cuda.synchronize(dev)
    if rank != src:
        eng.state_commit()
    return eng


def gather_shards(local: np.ndarray, M: int, group=None):
    """All-gather block-partitioned host results back into query order (optional; results normally stay sharded)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local), group=group)
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == M
    return out

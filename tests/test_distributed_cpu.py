"""CPU, world_size 2, gloo: the host-side logic of the query-sharded multi-GPU path -- block partition of the queries,
in-place broadcast of the model-state buffers from the fitting rank, and re-assembly of sharded results.  (The
device kernels are not involved; the GPU path itself is covered by bench.py --gpus N and tests/test_gpu_parity.py.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussian_process_transportation_b200.distributed import broadcast_tensors, gather_shards, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, M, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # model state: header, X, alpha, inverse factor (stand-ins with the real shapes for N=300, d=p=3, Npad=384)
        shapes = [32, 3 * 384, 3 * 384, 384 * 384]
        rng = np.random.default_rng(7)
        ref = [torch.from_numpy(rng.standard_normal(n)) for n in shapes]
        bufs = [r.clone() if rank == 0 else torch.zeros(n, dtype=torch.float64) for r, n in zip(ref, shapes)]
        broadcast_tensors(bufs, src=0)
        for b, r in zip(bufs, ref):
            assert torch.equal(b, r)
        # sharded "query": every rank evaluates f on its block, results are re-assembled in order
        x = np.arange(M, dtype=np.float64)[:, None] * np.ones((1, 3))
        lo, hi = shard_bounds(M, world, rank)
        local = np.sin(x[lo:hi]) * float(bufs[0][0])
        full = gather_shards(local, M)
        assert np.array_equal(full, np.sin(x) * float(ref[0][0]))
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([lo, hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("M", [10, 101])
def test_two_rank_sharding_and_broadcast(tmp_path, M):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, M, str(tmp_path)), nprocs=world, join=True)
    spans = [np.load(tmp_path / f"ok{r}.npy") for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == M and spans[0][1] == spans[1][0]


def test_shard_bounds_cover_everything():
    for M in [0, 1, 7, 64, 1000003]:
        for world in [1, 2, 4, 8]:
            spans = [shard_bounds(M, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1

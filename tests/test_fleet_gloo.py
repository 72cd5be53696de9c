"""Host-side logic of the multi-GPU path on CPU: world_size-2 and -3 process groups over gloo.
The solve itself needs a GPU; here a deterministic stand-in produces each rank's twists so that the sharding,
padding and the single all-gather can be checked for shard-invariance (the gathered tensor must not depend on the
number of ranks)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neo_mpc_planner2_b200.fleet import shard_bounds, max_shard, gather_twists


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_twists(lo, hi):
    idx = torch.arange(lo, hi, dtype=torch.float32)
    return torch.stack([idx * 0.5, -idx, idx * idx * 1e-3], dim=1)


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_total, world, rank)
        full = gather_twists(_fake_twists(lo, hi), n_total)
        q.put((rank, full.numpy().tobytes()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 1000), (2, 1001), (3, 1000)])
def test_gather_is_shard_invariant(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_twists(0, n_total).numpy().tobytes()
    for rank, blob in got:
        assert blob == want, f"rank {rank} gathered a different tensor"


def test_shard_bounds_cover_everything():
    """Contiguous cover, every shard at most ceil(n / world) rows — the library's neompc_shard_rows — so that the gathered
    rows of all ranks are contiguous; only trailing shards may be short or empty."""
    import ctypes
    from neo_mpc_planner2_b200 import _lib
    lib = _lib.load()
    for n in (0, 1, 7, 4096, 65537, 1048576):
        for world in (1, 2, 3, 4, 8):
            m = max_shard(n, world)
            assert m == lib.neompc_shard_rows(ctypes.c_size_t(n), world)
            edges = [shard_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) <= m and all(a == r * m or a == n for r, (a, b) in enumerate(edges))
            full = [sz for sz in sizes if sz == m]
            assert sizes[:len(full)] == full            # full shards first, then at most one short one, then empty ones

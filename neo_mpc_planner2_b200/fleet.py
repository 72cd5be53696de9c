"""Multi-GPU sharding of a batch of MPC problems: one process per GPU (torch.distributed), contiguous block
split of the request array, NO collective on the data path except ONE all-gather of the solved
(vx, vy, omega) at the end (BASELINE.json north_star).  Problems are independent — the reference
solves them one at a time (srv.py:349-403) — so results do not depend on the number of shards.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int, rank: int):
    """Rank r owns requests [r*n//world, (r+1)*n//world)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def max_shard(n: int, world: int) -> int:
    return max(shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world))


def gather_twists(local_twist, n_total: int, group=None):
    """All-gather the per-rank [n_local, 3] float32 twist tensors into the full [n_total, 3] tensor (every rank
    gets it).  Shards may differ by one row: they are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(n_total, world, rank)
    assert local_twist.shape == (hi - lo, 3), (local_twist.shape, hi - lo)
    m = max_shard(n_total, world)
    if hi - lo == m:
        send = local_twist.contiguous()
    else:
        send = torch.zeros((m, 3), dtype=local_twist.dtype, device=local_twist.device)
        send[: hi - lo] = local_twist
    out = torch.empty((world * m, 3), dtype=local_twist.dtype, device=local_twist.device)
    dist.all_gather_into_tensor(out, send, group=group)
    if world * m == n_total:
        return out
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_total, world, r)
        parts.append(out[r * m: r * m + (b - a)])
    return torch.cat(parts, dim=0)


class FleetSolver:
    """Solves a global batch across the ranks of a process group.  Every rank constructs it with the same
    parameters / costmap / footprint (replicated, <= 4 MB) and calls ``solve(all_requests)``."""

    def __init__(self, params, device: int, group=None, **over):
        import torch
        from .solver import BatchSolver
        self.torch = torch
        self.group = group
        self.device = torch.device("cuda", device)
        self.solver = BatchSolver(params, device=device, **over)
        self._bufs = None

    def load_workload(self, wl):
        self.solver.load_workload(wl)

    def solve(self, all_requests: np.ndarray):
        """all_requests: the full REQUEST_DTYPE array (same on every rank).  Returns (twist_all [B,3] torch tensor on
        the device, local responses as numpy)."""
        import torch.distributed as dist
        from .abi import REQUEST_DTYPE, RESPONSE_DTYPE
        torch = self.torch
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n_total = len(all_requests)
        lo, hi = shard_bounds(n_total, world, rank)
        local = np.ascontiguousarray(all_requests[lo:hi])
        n = hi - lo
        d_req = torch.from_numpy(local.view(np.uint8).reshape(n, REQUEST_DTYPE.itemsize)).to(self.device)
        d_out = torch.empty((n, RESPONSE_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        d_twist = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device)
        self.solver.solve_device(d_req.data_ptr(), n, d_out.data_ptr(), d_twist.data_ptr(), None, stream.cuda_stream)
        twist_all = gather_twists(d_twist, n_total, self.group) if world > 1 else d_twist
        resp = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=RESPONSE_DTYPE)
        return twist_all, resp

    def close(self):
        self.solver.close()

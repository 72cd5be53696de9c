"""Multi-GPU sharding of a batch of MPC problems (BASELINE.json north_star): contiguous block split of
the request array, NO collective on the data path except ONE all-gather of the solved (vx, vy, omega) at the end.  Problems
are independent — the reference solves them one at a time (srv.py:349-403) — so results do not depend on the number of
shards.  The sharding arithmetic and the collective are the library's (include/neompc.h: neompc_shard_rows,
neompc_solve_gather_device — NCCL, enqueued by the library behind its solve kernel); this module is the Python host side:

  FleetSolver   one process per GPU (torchrun): the NCCL id is created on rank 0 and distributed with torch.distributed
  LocalFleet    one process driving several GPUs (neompc_comm_init_all + neompc_fleet_solve): what a C++ fleet host does
"""
from __future__ import annotations

import ctypes

import numpy as np


def shard_rows(n: int, world: int) -> int:
    """neompc_shard_rows: ceil(n / world)."""
    return (n + world - 1) // world


def shard_bounds(n: int, world: int, rank: int):
    """Rank r owns requests [r*m, (r+1)*m) clipped to n, m = ceil(n / world): the gathered rows are then contiguous."""
    m = shard_rows(n, world)
    return min(n, rank * m), min(n, (rank + 1) * m)


def max_shard(n: int, world: int) -> int:
    return shard_rows(n, world)


def gather_twists(local_twist, n_total: int, group=None):
    """The layout of the collective restated with torch.distributed (any backend; used by the CPU tests over gloo): every
    rank contributes shard_rows(n, world) rows (a short last shard is zero-padded), all ranks receive [n_total, 3]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(n_total, world, rank)
    assert local_twist.shape == (hi - lo, 3), (local_twist.shape, hi - lo)
    m = shard_rows(n_total, world)
    send = torch.zeros((m, 3), dtype=local_twist.dtype, device=local_twist.device)
    send[: hi - lo] = local_twist
    out = torch.empty((world * m, 3), dtype=local_twist.dtype, device=local_twist.device)
    dist.all_gather_into_tensor(out, send, group=group)
    return out[:n_total]


class FleetSolver:
    """One rank of a fleet: every rank constructs it with the same parameters / costmap / footprint (replicated) and
    calls ``solve(all_requests)``; the all-gather runs inside libneompc (NCCL)."""

    def __init__(self, params, device: int, group=None, **over):
        import torch
        import torch.distributed as dist
        from .solver import BatchSolver
        self.torch = torch
        self.group = group
        self.device = torch.device("cuda", device)
        self.solver = BatchSolver(params, device=device, **over)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        # the NCCL id of the library's own communicator travels over the process group that already exists
        backend = dist.get_backend(group) if dist.is_initialized() else None
        dev = self.device if backend == "nccl" else torch.device("cpu")
        if self.rank == 0:
            uid = torch.frombuffer(bytearray(BatchSolver.comm_unique_id()), dtype=torch.uint8).to(dev)
        else:
            uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if self.world > 1:
            dist.broadcast(uid, src=0, group=group)
        self.solver.comm_init(bytes(uid.cpu().numpy().tobytes()), self.world, self.rank)

    def load_workload(self, wl):
        self.solver.load_workload(wl)

    def solve_gather_device(self, d_reqs, n_local, rows, d_out, d_twist_all, stream):
        """Device tensors in, asynchronous on `stream` (torch.cuda.Stream): see neompc_solve_gather_device."""
        self.solver.solve_gather_device(d_reqs.data_ptr(), n_local, rows, d_out.data_ptr(), d_twist_all.data_ptr(),
                                        stream.cuda_stream)

    def solve(self, all_requests: np.ndarray):
        """all_requests: the full REQUEST_DTYPE array (same on every rank).  Returns (twist_all [B,3] torch tensor on
        the device, this rank's responses as numpy)."""
        from .abi import REQUEST_DTYPE, RESPONSE_DTYPE
        torch = self.torch
        n_total = len(all_requests)
        lo, hi = shard_bounds(n_total, self.world, self.rank)
        rows = shard_rows(n_total, self.world)
        local = np.ascontiguousarray(all_requests[lo:hi])
        n = hi - lo
        d_req = torch.from_numpy(local.view(np.uint8).reshape(n, REQUEST_DTYPE.itemsize)).to(self.device)
        d_out = torch.empty((max(n, 1), RESPONSE_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        d_all = torch.empty((self.world * rows, 3), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device)
        self.solve_gather_device(d_req, n, rows, d_out, d_all, stream)
        self.solver.gather_wait(stream.cuda_stream, 0)
        stream.synchronize()
        resp = np.frombuffer(d_out[:n].cpu().numpy().tobytes(), dtype=RESPONSE_DTYPE)
        return d_all[:n_total], resp

    def close(self):
        self.solver.close()


class LocalFleet:
    """One process, several GPUs: handles[i] is rank i of one communicator (neompc_comm_init_all); ``solve`` is
    neompc_fleet_solve — host requests in, all twists (and optionally all responses) out."""

    def __init__(self, params, devices, **over):
        from . import _lib
        from .solver import BatchSolver
        self._lib = _lib.load()
        _lib.prefer_torch_nccl()
        self.solvers = [BatchSolver(params, device=d, **over) for d in devices]
        self._arr = (ctypes.c_void_p * len(self.solvers))(*[s._h for s in self.solvers])
        rc = self._lib.neompc_comm_init_all(self._arr, len(self.solvers))
        if rc != 0:
            msg = self._lib.neompc_last_error(self.solvers[0]._h).decode()
            self.close()
            raise _lib.NeompcError(f"neompc_comm_init_all failed ({rc}): {msg}")

    def load_workload(self, wl):
        for s in self.solvers:
            s.load_workload(wl)

    def reserve_instances(self, n):
        for s in self.solvers:
            s.reserve_instances(n)

    def solve(self, reqs, want_responses=False):
        from .abi import REQUEST_DTYPE, RESPONSE_DTYPE
        reqs = np.ascontiguousarray(reqs, dtype=REQUEST_DTYPE)
        n = len(reqs)
        twist = np.empty((n, 3), np.float32)
        out = np.empty(n, RESPONSE_DTYPE) if want_responses else None
        rc = self._lib.neompc_fleet_solve(self._arr, len(self.solvers), reqs.ctypes.data_as(ctypes.c_void_p), n,
                                          twist.ctypes.data_as(ctypes.c_void_p),
                                          out.ctypes.data_as(ctypes.c_void_p) if out is not None else None)
        if rc != 0:
            from ._lib import NeompcError
            raise NeompcError(f"neompc_fleet_solve failed ({rc}): {self._lib.neompc_last_error(self.solvers[0]._h).decode()}")
        return (twist, out) if want_responses else twist

    def gathered_on(self, rank, n):
        """Rank `rank`'s copy of the last gather (all copies are identical)."""
        tw = np.empty((n, 3), np.float32)
        s = self.solvers[rank]
        s._check(self._lib.neompc_fleet_get_gathered(s._h, n, tw.ctypes.data_as(ctypes.c_void_p)), "neompc_fleet_get_gathered")
        return tw

    def close(self):
        for s in getattr(self, "solvers", []):
            s.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

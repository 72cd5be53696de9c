"""BatchSolver — thin Python handle on libneompc (the C ABI does the work; see include/neompc.h).

Host-side mirror of what the reference's server object owns: parameters (srv.py:49-103), the costmap
(srv.py:118), the footprint (srv.py:154-155) and the per-instance state of ``optimizer()`` (srv.py:349-403).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .abi import (REQUEST_DTYPE, RESPONSE_DTYPE, PARAMS_DTYPE, MSG_DTYPE, ENC_OCCUPANCY, params_record,
                  TICK_DTYPE, CARROT_INFO_DTYPE, CARROT_PARAMS_DTYPE, PLAN_POSE_DTYPE, STATELESS)

NeompcError = _lib.NeompcError


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class BatchSolver:
    def __init__(self, params=None, device: int = 0, **over):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self.params = params_record(params, **over)
        rc = self._lib.neompc_create(_ptr(self.params), int(device), ctypes.byref(self._h))
        if rc != 0:
            msg = self._lib.neompc_last_error(None).decode()
            self._h = ctypes.c_void_p()
            raise NeompcError(f"neompc_create failed ({rc}): {msg}")
        self.device = int(device)
        self.control_steps = int(self.params["control_steps"])

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc, what):
        if rc != 0:
            raise NeompcError(f"{what} failed ({rc}): {self._lib.neompc_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.neompc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ environment
    def set_params(self, params=None, **over):
        """neompc_set_params.  Without `params` the CURRENT record is the starting point: every field that is not named —
        solver knobs included — keeps its value."""
        rec = params_record(self.params if params is None else params, **over)
        self._check(self._lib.neompc_set_params(self._h, _ptr(rec)), "neompc_set_params")
        self.params = rec
        self.control_steps = int(rec["control_steps"])

    def get_params(self):
        """The record the library holds (neompc_get_params)."""
        rec = np.zeros((), dtype=PARAMS_DTYPE)
        self._check(self._lib.neompc_get_params(self._h, _ptr(rec)), "neompc_get_params")
        return rec

    def set_costmap(self, cells, resolution, origin_x, origin_y, encoding=ENC_OCCUPANCY):
        if cells is None:
            self._check(self._lib.neompc_set_costmap(self._h, None, 0, 0, 1.0, 0.0, 0.0, encoding), "neompc_set_costmap")
            return
        cells = np.ascontiguousarray(cells)
        if cells.dtype == np.int8:
            cells = cells.view(np.uint8)
        if cells.dtype != np.uint8 or cells.ndim != 2:
            raise ValueError("cells must be uint8/int8 [H, W]")
        h, w = cells.shape
        self._check(self._lib.neompc_set_costmap(self._h, _ptr(cells), w, h, float(resolution), float(origin_x),
                                                 float(origin_y), int(encoding)), "neompc_set_costmap")

    def set_costmap_device(self, cells_tensor, resolution, origin_x, origin_y, encoding=ENC_OCCUPANCY):
        h, w = cells_tensor.shape
        self._check(self._lib.neompc_set_costmap_device(self._h, ctypes.c_void_p(cells_tensor.data_ptr()), w, h,
                                                        float(resolution), float(origin_x), float(origin_y),
                                                        int(encoding)), "neompc_set_costmap_device")

    def set_footprint(self, xy):
        xy = np.ascontiguousarray(np.asarray(xy, dtype=np.float32).reshape(-1, 2))
        self._check(self._lib.neompc_set_footprint(self._h, _ptr(xy), len(xy)), "neompc_set_footprint")

    def load_workload(self, wl):
        if wl.cells is not None:
            self.set_costmap(wl.cells, wl.resolution, wl.origin_x, wl.origin_y, wl.encoding)
        else:
            self.set_costmap(None, 1.0, 0.0, 0.0)
        self.set_footprint(wl.footprint)

    # ------------------------------------------------------------------ per-instance state
    def reserve_instances(self, n):
        self._check(self._lib.neompc_reserve_instances(self._h, int(n)), "neompc_reserve_instances")

    def reset_state(self, ids=None):
        if ids is None:
            self._check(self._lib.neompc_reset_state(self._h, None, 0), "neompc_reset_state")
        else:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            self._check(self._lib.neompc_reset_state(self._h, _ptr(ids), len(ids)), "neompc_reset_state")

    def get_state(self, instance_id):
        guess = np.zeros(3 * self.control_steps, np.float32)
        last = np.zeros(3, np.float32)
        wait = ctypes.c_float()
        flags = ctypes.c_uint32()
        self._check(self._lib.neompc_get_state(self._h, int(instance_id), _ptr(guess), _ptr(last),
                                               ctypes.byref(wait), ctypes.byref(flags)), "neompc_get_state")
        return dict(initial_guess=guess, last_control=last, waiting_time=wait.value, flags=flags.value)

    # ------------------------------------------------------------------ the hot path
    def solve(self, reqs, want_plan=False, out=None, plan_out=None):
        """Host buffers in, host buffers out (neompc_solve_batch)."""
        reqs = np.ascontiguousarray(reqs, dtype=REQUEST_DTYPE)
        n = len(reqs)
        if out is None:
            out = np.empty(n, RESPONSE_DTYPE)
        plan = None
        if want_plan:
            plan = plan_out if plan_out is not None else np.empty((n, 3 * self.control_steps), np.float32)
        self._check(self._lib.neompc_solve_batch(self._h, _ptr(reqs), n, _ptr(out), _ptr(plan)), "neompc_solve_batch")
        return (out, plan) if want_plan else out

    def solve_raw(self, reqs_ptr, n, out_ptr, plan_ptr=None):
        """neompc_solve_batch on raw host addresses (e.g. pinned torch tensors)."""
        self._check(self._lib.neompc_solve_batch(self._h, ctypes.c_void_p(reqs_ptr), int(n), ctypes.c_void_p(out_ptr),
                                                 ctypes.c_void_p(plan_ptr) if plan_ptr else None), "neompc_solve_batch")

    def solve_twists(self, reqs, out=None):
        """neompc_solve_batch_twists: host requests in, [n, 3] (vx, vy, omega) out."""
        reqs = np.ascontiguousarray(reqs, dtype=REQUEST_DTYPE)
        n = len(reqs)
        if out is None:
            out = np.empty((n, 3), np.float32)
        self._check(self._lib.neompc_solve_batch_twists(self._h, _ptr(reqs), n, _ptr(out)), "neompc_solve_batch_twists")
        return out

    def solve_twists_raw(self, reqs_ptr, n, twist_ptr):
        self._check(self._lib.neompc_solve_batch_twists(self._h, ctypes.c_void_p(reqs_ptr), int(n), ctypes.c_void_p(twist_ptr)),
                    "neompc_solve_batch_twists")

    def solve_device(self, d_reqs, n, d_out, d_twist=None, d_plan=None, stream=None):
        """Device pointers (ints) in and out, asynchronous on `stream`: a cudaStream_t as int (0 = the legacy default
        stream, which CUDA names by the handle 0x1), or None for the handle's own stream."""
        if stream == 0:
            stream = 1          # cudaStreamLegacy
        self._check(self._lib.neompc_solve_batch_device(
            self._h, ctypes.c_void_p(d_reqs), int(n), ctypes.c_void_p(d_out),
            ctypes.c_void_p(d_twist) if d_twist else None, ctypes.c_void_p(d_plan) if d_plan else None,
            ctypes.c_void_p(stream) if stream else None), "neompc_solve_batch_device")

    # ------------------------------------------------------------------ multi-GPU: one rank of a fleet (include/neompc.h)
    @staticmethod
    def comm_unique_id():
        """128 opaque bytes rank 0 creates and the caller distributes (neompc_comm_unique_id)."""
        lib = _lib.load()
        _lib.prefer_torch_nccl()
        buf = (ctypes.c_ubyte * 128)()
        rc = lib.neompc_comm_unique_id(buf)
        if rc != 0:
            raise NeompcError(f"neompc_comm_unique_id failed ({rc}): {lib.neompc_last_error(None).decode()}")
        return bytes(buf)

    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int):
        _lib.prefer_torch_nccl()
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self._lib.neompc_comm_init(self._h, buf, int(n_ranks), int(rank)), "neompc_comm_init")

    def comm_info(self):
        n, r = ctypes.c_int(), ctypes.c_int()
        self._lib.neompc_comm_info(self._h, ctypes.byref(n), ctypes.byref(r))
        return n.value, r.value

    def solve_gather_device(self, d_reqs, n_local, shard_rows, d_out, d_twist_all, stream=None):
        """This rank's shard + the all-gather of every rank's (vx, vy, omega) into d_twist_all [n_ranks*shard_rows, 3]."""
        if stream == 0:
            stream = 1
        self._check(self._lib.neompc_solve_gather_device(
            self._h, ctypes.c_void_p(d_reqs), int(n_local), int(shard_rows), ctypes.c_void_p(d_out),
            ctypes.c_void_p(d_twist_all), ctypes.c_void_p(stream) if stream else None), "neompc_solve_gather_device")

    def gather_wait(self, stream=None, age=0):
        if stream == 0:
            stream = 1
        self._check(self._lib.neompc_gather_wait(self._h, ctypes.c_void_p(stream) if stream else None, int(age)),
                    "neompc_gather_wait")

    def solve_msgs(self, msgs, want_plan=False):
        msgs = np.ascontiguousarray(msgs, dtype=MSG_DTYPE)
        n = len(msgs)
        out = np.empty(n, RESPONSE_DTYPE)
        plan = np.empty((n, 3 * self.control_steps), np.float32) if want_plan else None
        self._check(self._lib.neompc_solve_msgs(self._h, _ptr(msgs), n, _ptr(out), _ptr(plan)), "neompc_solve_msgs")
        return (out, plan) if want_plan else out

    def pack_requests_device(self, d_msgs, n, d_reqs, stream=None):
        self._check(self._lib.neompc_pack_requests(self._h, ctypes.c_void_p(d_msgs), int(n), ctypes.c_void_p(d_reqs),
                                                   ctypes.c_void_p(stream) if stream else None), "neompc_pack_requests")

    # ------------------------------------------------------------------ the step before: carrot selection (cpp:66-246)
    def set_plan(self, xyyaw):
        """setPlan (cpp:274-281): global plan poses (x, y, yaw), shared by all robots of this handle."""
        plan = np.ascontiguousarray(np.asarray(xyyaw, dtype=np.float64).reshape(-1, 3))
        self._check(self._lib.neompc_set_plan(self._h, _ptr(plan), len(plan)), "neompc_set_plan")

    @staticmethod
    def carrot_params(lookahead_dist_min=0.5, lookahead_dist_max=0.5, lookahead_dist_close_to_goal=0.5,
                      controller_frequency=30.0):
        cp = np.zeros((), CARROT_PARAMS_DTYPE)
        cp["lookahead_dist_min"], cp["lookahead_dist_max"] = lookahead_dist_min, lookahead_dist_max
        cp["lookahead_dist_close_to_goal"], cp["controller_frequency"] = lookahead_dist_close_to_goal, controller_frequency
        return cp

    def build_requests(self, ticks, carrot_params, first_instance_id=STATELESS):
        """Carrot selection + request construction for n robots (neompc_build_requests, host buffers)."""
        ticks = np.ascontiguousarray(ticks, dtype=TICK_DTYPE)
        n = len(ticks)
        reqs = np.empty(n, REQUEST_DTYPE)
        info = np.empty(n, CARROT_INFO_DTYPE)
        self._check(self._lib.neompc_build_requests(self._h, _ptr(carrot_params), _ptr(ticks), n,
                                                    int(first_instance_id), _ptr(reqs), _ptr(info)),
                    "neompc_build_requests")
        return reqs, info

    def control_tick(self, ticks, carrot_params, first_instance_id=STATELESS, want_plan=False, want_requests=False):
        """neompc_control_tick: carrot selection + solve for n robots in one call.  Returns (responses, carrot info[, plan]
        [, requests])."""
        ticks = np.ascontiguousarray(ticks, dtype=TICK_DTYPE)
        n = len(ticks)
        out = np.empty(n, RESPONSE_DTYPE)
        info = np.empty(n, CARROT_INFO_DTYPE)
        plan = np.empty((n, 3 * self.control_steps), np.float32) if want_plan else None
        reqs = np.empty(n, REQUEST_DTYPE) if want_requests else None
        rc = self._lib.neompc_control_tick(self._h, _ptr(carrot_params), _ptr(ticks), n, int(first_instance_id), _ptr(out),
                                           _ptr(info), _ptr(reqs), _ptr(plan))
        self._check(rc, "neompc_control_tick")
        res = [out, info]
        if want_plan:
            res.append(plan)
        if want_requests:
            res.append(reqs)
        return tuple(res)

    def build_requests_device(self, carrot_params, d_ticks, n, d_reqs, d_info, first_instance_id=STATELESS, stream=None):
        if stream == 0:
            stream = 1
        self._check(self._lib.neompc_build_requests_device(
            self._h, _ptr(carrot_params), ctypes.c_void_p(d_ticks), int(n), int(first_instance_id),
            ctypes.c_void_p(d_reqs), ctypes.c_void_p(d_info), ctypes.c_void_p(stream) if stream else None),
            "neompc_build_requests_device")

    # ------------------------------------------------------------------ the step after: predicted path (srv.py:271-310)
    def local_plan(self, reqs, plan):
        """publishLocalPlan for n solved problems: [n, control_steps + 1] poses (x, y, qz, qw), host buffers."""
        reqs = np.ascontiguousarray(reqs, dtype=REQUEST_DTYPE)
        n = len(reqs)
        plan = np.ascontiguousarray(plan, dtype=np.float32).reshape(n, 3 * self.control_steps)
        poses = np.empty((n, self.control_steps + 1), PLAN_POSE_DTYPE)
        self._check(self._lib.neompc_local_plan(self._h, _ptr(reqs), _ptr(plan), n, _ptr(poses)), "neompc_local_plan")
        return poses

    def local_plan_device(self, d_reqs, d_plan, n, d_poses, stream=None):
        if stream == 0:
            stream = 1
        self._check(self._lib.neompc_local_plan_device(
            self._h, ctypes.c_void_p(d_reqs), ctypes.c_void_p(d_plan), int(n), ctypes.c_void_p(d_poses),
            ctypes.c_void_p(stream) if stream else None), "neompc_local_plan_device")

    def eval_objective(self, reqs, u, want_grad=True):
        reqs = np.ascontiguousarray(reqs, dtype=REQUEST_DTYPE)
        n = len(reqs)
        u = np.ascontiguousarray(u, dtype=np.float32).reshape(n, 3 * self.control_steps)
        J = np.empty(n, np.float32)
        g = np.empty_like(u) if want_grad else None
        self._check(self._lib.neompc_eval_objective(self._h, _ptr(reqs), _ptr(u), n, _ptr(J), _ptr(g)),
                    "neompc_eval_objective")
        return (J, g) if want_grad else J

    @property
    def launch_count(self):
        return int(self._lib.neompc_launch_count(self._h))

    def tiling_for(self, n):
        """(lanes per instance, steps per lane) the dispatcher uses for a batch of n requests (latency tiling while the batch
        is resident at once, else the throughput tiling of `tiling`)."""
        g, s = ctypes.c_int(), ctypes.c_int()
        self._lib.neompc_get_tiling_for(self._h, int(n), ctypes.byref(g), ctypes.byref(s))
        return g.value, s.value

    @property
    def last_host_path(self):
        """neompc_last_host_path: 1 mailbox, 2 chunked staged copies, 3 zero-copy (pinned, device-accessible buffers)."""
        return int(self._lib.neompc_last_host_path(self._h))

    @property
    def tiling(self):
        g, s = ctypes.c_int(), ctypes.c_int()
        self._lib.neompc_get_tiling(self._h, ctypes.byref(g), ctypes.byref(s))
        return g.value, s.value

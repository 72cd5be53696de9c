"""ctypes binding of libneompc.so (include/neompc.h).  Fails loudly when the library is missing:
there is no CPU fallback for the solve."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NEOMPC_LIB selects another build of the same library (tuning variants made by scripts/build_variant.sh)
LIB_PATH = os.environ.get("NEOMPC_LIB") or os.path.join(HERE, "libneompc.so")

# every symbol include/neompc.h declares
EXPORTS = [
    "neompc_create", "neompc_destroy", "neompc_set_params", "neompc_get_params", "neompc_last_error",
    "neompc_version", "neompc_abi_sizes", "neompc_set_costmap", "neompc_set_costmap_device",
    "neompc_set_footprint", "neompc_reserve_instances", "neompc_reset_state", "neompc_get_state",
    "neompc_solve_batch", "neompc_solve_batch_device", "neompc_solve_msgs", "neompc_pack_requests",
    "neompc_set_plan", "neompc_build_requests", "neompc_build_requests_device",
    "neompc_local_plan", "neompc_local_plan_device",
    "neompc_eval_objective", "neompc_launch_count", "neompc_last_host_path", "neompc_get_tiling", "neompc_get_tiling_for", "neompc_host_alloc", "neompc_host_free",
    "neompc_comm_unique_id", "neompc_comm_init", "neompc_comm_init_all", "neompc_comm_destroy", "neompc_comm_info",
    "neompc_shard_rows", "neompc_solve_gather_device", "neompc_gather_wait", "neompc_fleet_solve",
    "neompc_fleet_get_gathered", "neompc_control_tick", "neompc_solve_batch_twists",
]

_lib = None


class NeompcError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libneompc.so in-tree with nvcc for sm_100a (make -C csrc)."""
    import subprocess
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise NeompcError("building libneompc.so failed")
    return LIB_PATH


def prefer_torch_nccl():
    """libneompc binds NCCL at run time and takes the copy the process has loaded already.  In a Python process that copy
    should be torch's (bundled, same SONAME as the system's): binding the system copy first would break a later
    `import torch`.  Called before the first communicator call."""
    try:
        import torch  # noqa: F401
        import torch.distributed  # noqa: F401
    except Exception:  # no torch: the system's libnccl.so.2 is used
        pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NeompcError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or make -C neo_mpc_planner2_b200/csrc).  neo_mpc_planner2_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, u32, sz, i32, f64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_int, ctypes.c_double
    lib.neompc_create.argtypes = [vp, i32, ctypes.POINTER(vp)]
    lib.neompc_destroy.argtypes = [vp]
    lib.neompc_set_params.argtypes = [vp, vp]
    lib.neompc_get_params.argtypes = [vp, vp]
    lib.neompc_last_error.argtypes = [vp]
    lib.neompc_last_error.restype = ctypes.c_char_p
    lib.neompc_version.restype = i32
    lib.neompc_abi_sizes.argtypes = [ctypes.POINTER(sz)]
    lib.neompc_set_costmap.argtypes = [vp, vp, u32, u32, f64, f64, f64, i32]
    lib.neompc_set_costmap_device.argtypes = [vp, vp, u32, u32, f64, f64, f64, i32]
    lib.neompc_set_footprint.argtypes = [vp, vp, i32]
    lib.neompc_reserve_instances.argtypes = [vp, u32]
    lib.neompc_reset_state.argtypes = [vp, vp, sz]
    lib.neompc_get_state.argtypes = [vp, u32, vp, vp, vp, vp]
    lib.neompc_solve_batch.argtypes = [vp, vp, sz, vp, vp]
    lib.neompc_solve_batch_twists.argtypes = [vp, vp, sz, vp]
    lib.neompc_solve_batch_device.argtypes = [vp, vp, sz, vp, vp, vp, vp]
    lib.neompc_solve_msgs.argtypes = [vp, vp, sz, vp, vp]
    lib.neompc_pack_requests.argtypes = [vp, vp, sz, vp, vp]
    lib.neompc_set_plan.argtypes = [vp, vp, sz]
    lib.neompc_build_requests.argtypes = [vp, vp, vp, sz, u32, vp, vp]
    lib.neompc_build_requests_device.argtypes = [vp, vp, vp, sz, u32, vp, vp, vp]
    lib.neompc_local_plan.argtypes = [vp, vp, vp, sz, vp]
    lib.neompc_local_plan_device.argtypes = [vp, vp, vp, sz, vp, vp]
    lib.neompc_eval_objective.argtypes = [vp, vp, vp, sz, vp, vp]
    lib.neompc_launch_count.argtypes = [vp]
    lib.neompc_launch_count.restype = ctypes.c_uint64
    lib.neompc_get_tiling_for.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    lib.neompc_get_tiling_for.restype = ctypes.c_int
    lib.neompc_last_host_path.argtypes = [vp]
    lib.neompc_last_host_path.restype = ctypes.c_int
    lib.neompc_get_tiling.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    lib.neompc_host_alloc.argtypes = [ctypes.POINTER(vp), sz]
    lib.neompc_host_free.argtypes = [vp]
    lib.neompc_comm_unique_id.argtypes = [vp]
    lib.neompc_comm_init.argtypes = [vp, vp, i32, i32]
    lib.neompc_comm_init_all.argtypes = [ctypes.POINTER(vp), i32]
    lib.neompc_comm_destroy.argtypes = [vp]
    lib.neompc_comm_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    lib.neompc_shard_rows.argtypes = [sz, i32]
    lib.neompc_shard_rows.restype = sz
    lib.neompc_solve_gather_device.argtypes = [vp, vp, sz, sz, vp, vp, vp]
    lib.neompc_gather_wait.argtypes = [vp, vp, i32]
    lib.neompc_fleet_solve.argtypes = [ctypes.POINTER(vp), i32, vp, sz, vp, vp]
    lib.neompc_fleet_get_gathered.argtypes = [vp, sz, vp]
    lib.neompc_control_tick.argtypes = [vp, vp, vp, sz, u32, vp, vp, vp, vp]
    _lib = lib
    return lib

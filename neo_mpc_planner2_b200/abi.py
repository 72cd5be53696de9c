"""numpy mirrors of the POD records declared in ``include/neompc.h``.

Field order, sizes and offsets must match the C structs exactly; ``tests/test_abi.py`` checks the
sizes against ``neompc_abi_sizes()`` exported by the library.
"""
from __future__ import annotations

import numpy as np

STATELESS = 0xFFFFFFFF          # neompc_request.instance_id: cold start, no per-instance state

# neompc_request — mirror of neo_srvs2/srv/Optimizer.Request (reference cpp:240-246,
# srv.py:350-355), planar form, 64 bytes
REQUEST_DTYPE = np.dtype([
    ("vel_x", "<f4"), ("vel_y", "<f4"), ("vel_theta", "<f4"),
    ("carrot_x", "<f4"), ("carrot_y", "<f4"), ("carrot_yaw", "<f4"),
    ("goal_x", "<f4"), ("goal_y", "<f4"), ("goal_yaw", "<f4"),
    ("pose_x", "<f4"), ("pose_y", "<f4"), ("pose_yaw", "<f4"),
    ("pose_yaw_objective", "<f4"),
    ("control_interval", "<f4"),
    ("delta_t", "<f4"),
    ("instance_id", "<u4"),
], align=False)
assert REQUEST_DTYPE.itemsize == 64

# neompc_response — Optimizer.Response.output_vel (srv.py:375-377,389-391) + solver diagnostics, 32 bytes
RESPONSE_DTYPE = np.dtype([
    ("vx", "<f4"), ("vy", "<f4"), ("omega", "<f4"),
    ("cost", "<f4"),
    ("iters", "<u4"), ("evals", "<u4"),
    ("status", "<u4"), ("flags", "<u4"),
], align=False)
assert RESPONSE_DTYPE.itemsize == 32

# neompc_optimizer_request — float64 mirror of neo_srvs2/srv/Optimizer.Request with quaternions, 240 bytes
MSG_DTYPE = np.dtype([
    ("current_vel", "<f8", (6,)),      # Twist: linear xyz, angular xyz
    ("carrot_pose", "<f8", (7,)),      # position xyz, orientation xyzw
    ("goal_pose", "<f8", (7,)),
    ("current_pose", "<f8", (7,)),
    ("control_interval", "<f8"),
    ("delta_t", "<f8"),
    ("instance_id", "<u4"),
    ("switch_opt", "<u4"),
], align=False)
assert MSG_DTYPE.itemsize == 240

# neompc_robot_tick / neompc_carrot_info / neompc_carrot_params — carrot selection (cpp:66-246)
TICK_DTYPE = np.dtype([
    ("pose_x", "<f8"), ("pose_y", "<f8"), ("pose_yaw", "<f8"),
    ("vel_x", "<f4"), ("vel_y", "<f4"), ("vel_theta", "<f4"),
    ("plan_start", "<u4"), ("slow_down", "<u4"), ("delta_t", "<f4"),
], align=False)
assert TICK_DTYPE.itemsize == 48
CARROT_INFO_DTYPE = np.dtype([("status", "<u4"), ("plan_start", "<u4"), ("carrot_index", "<u4"), ("flags", "<u4")])
assert CARROT_INFO_DTYPE.itemsize == 16
CARROT_PARAMS_DTYPE = np.dtype([("lookahead_dist_min", "<f4"), ("lookahead_dist_max", "<f4"),
                                ("lookahead_dist_close_to_goal", "<f4"), ("controller_frequency", "<f4")])

# neompc_plan_pose — one pose of the predicted path (publishLocalPlan, srv.py:271-310), 32 bytes
PLAN_POSE_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("qz", "<f8"), ("qw", "<f8")])
assert PLAN_POSE_DTYPE.itemsize == 32

# neompc_response.status
STATUS_CONVERGED = 0
STATUS_MAXITER = 1
STATUS_LINESEARCH = 2

# neompc_response.flags
FLAG_COLLISION = 1            # self.collision latched (srv.py:338-339)
FLAG_COLLISION_FOOTPRINT = 2  # self.collision_footprint (srv.py:343-347)
FLAG_NEW_GOAL = 4             # the new-goal reset ran (srv.py:358-361)
FLAG_STOPPED = 8              # zero twist returned (srv.py:374-377)

ENC_OCCUPANCY = 0
ENC_NAV2_RAW = 1

# neompc_params.footprint_mode
FOOTPRINT_STATIC = 0          # the reference: the polygon never moves (aliasing at srv.py:227,241-244)
FOOTPRINT_MOVING = 1          # opt-in: polygon placed at every predicted pose

# neompc_params.costmap_mode
COSTMAP_NEAREST = 0           # the reference: cost of the cell under the predicted position
COSTMAP_BILINEAR = 1          # opt-in: bilinear interpolation between cell centres, gradient enters the solver (row N4)

# neompc_params.costmap_guidance (solver strategy; the objective stays the reference's)
GUIDANCE_ON = 0
GUIDANCE_OFF = 1

# neompc_params — the reference's 22 server parameters (srv.py:49-75) + solver knobs
PARAMS_FIELDS = [
    ("acc_x_limit", "<f4"), ("acc_y_limit", "<f4"), ("acc_theta_limit", "<f4"),
    ("min_vel_x", "<f4"), ("min_vel_y", "<f4"), ("min_vel_trans", "<f4"), ("min_vel_theta", "<f4"),
    ("max_vel_x", "<f4"), ("max_vel_y", "<f4"), ("max_vel_trans", "<f4"), ("max_vel_theta", "<f4"),
    ("w_trans", "<f4"), ("w_orient", "<f4"), ("w_control", "<f4"), ("w_terminal", "<f4"),
    ("w_costmap", "<f4"), ("w_footprint", "<f4"),
    ("waiting_time", "<f4"), ("low_pass_gain", "<f4"), ("opt_tolerance", "<f4"),
    ("prediction_horizon", "<f4"),
    ("control_steps", "<i4"),
    # solver knobs (not in the reference; 0 selects the library default)
    ("max_iterations", "<i4"),
    ("lbfgs_memory", "<i4"),
    ("control_smoothing", "<f4"),
    ("lanes_per_instance", "<i4"),
    ("footprint_mode", "<i4"),
    ("costmap_mode", "<i4"),
    ("costmap_guidance", "<i4"),
    ("reserved", "<i4", (3,)),
]
PARAMS_DTYPE = np.dtype(PARAMS_FIELDS, align=False)
assert PARAMS_DTYPE.itemsize == 128

REFERENCE_PARAM_NAMES = [f[0] for f in PARAMS_FIELDS[:22]]
KNOB_NAMES = [f[0] for f in PARAMS_FIELDS[22:] if f[0] != "reserved"]


def params_record(params=None, **over) -> np.ndarray:
    """Build a ``neompc_params`` record from an object/dict with the reference's parameter names
    (defaults: the code defaults of srv.py:49-75)."""
    defaults = dict(
        acc_x_limit=0.5, acc_y_limit=0.5, acc_theta_limit=0.5,
        min_vel_x=-0.5, min_vel_y=-0.5, min_vel_trans=0.5, min_vel_theta=-0.5,
        max_vel_x=0.5, max_vel_y=0.5, max_vel_trans=0.5, max_vel_theta=0.5,
        w_trans=0.5, w_orient=0.5, w_control=0.5, w_terminal=0.5, w_costmap=0.5, w_footprint=2000,
        waiting_time=3.0, low_pass_gain=0.5, opt_tolerance=1e-5, prediction_horizon=0.5,
        control_steps=3)
    rec = np.zeros((), dtype=PARAMS_DTYPE)
    src = {}
    if params is not None:
        if isinstance(params, dict):
            src = params
        elif isinstance(params, np.ndarray) and params.dtype.names:
            src = {k: params[k].item() for k in params.dtype.names if k != "reserved"}
        else:
            src = {k: getattr(params, k) for k in KNOB_NAMES + REFERENCE_PARAM_NAMES if hasattr(params, k)}
    unknown = [k for k in list(src) + list(over) if k not in REFERENCE_PARAM_NAMES and k not in KNOB_NAMES]
    if unknown:
        raise KeyError(f"unknown neompc_params field(s): {unknown}")
    for name in REFERENCE_PARAM_NAMES:
        rec[name] = over.pop(name, src.get(name, defaults[name]))
    for name in KNOB_NAMES:                      # solver knobs travel with the record (0 = library default)
        rec[name] = over.pop(name, src.get(name, 0))
    return rec


README_SAMPLE = dict(
    acc_x_limit=2.5, acc_y_limit=2.5, acc_theta_limit=3.0,
    min_vel_x=-0.7, min_vel_y=-0.7, min_vel_trans=-0.7, min_vel_theta=-0.7,
    max_vel_x=0.7, max_vel_y=0.7, max_vel_trans=0.7, max_vel_theta=0.7,
    w_trans=0.82, w_orient=0.50, w_control=0.05, w_terminal=0.05,
    w_footprint=0, w_costmap=0.05, waiting_time=3.0, low_pass_gain=0.5,
    opt_tolerance=1e-3, prediction_horizon=0.8, control_steps=3)

"""CPU restatement of the reference hot path (ORACLE / test infrastructure only).

Every function cites the lines of ``/root/reference/neo_mpc_planner2/mpc_optimization_server.py``
("srv.py") it follows.  The scalar functions are written to be BIT-IDENTICAL in float64 to the
reference (same numpy primitives in the same order); ``tests/golden/make_golden.py`` checks that
against the unmodified reference file and freezes the outputs in ``tests/golden/*.json``.

Problems are described by records with the field names of ``neompc_request``
(``include/neompc.h``): vel_x, vel_y, vel_theta, carrot_x, carrot_y, carrot_yaw, goal_x, goal_y,
goal_yaw, pose_x, pose_y, pose_yaw, pose_yaw_objective, control_interval, delta_t, instance_id.
Anything with those attributes/keys works (numpy structured scalar, dict, ``Problem``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from functools import partial

import numpy as np
from scipy.optimize import minimize

REQUEST_FIELDS = (
    "vel_x", "vel_y", "vel_theta", "carrot_x", "carrot_y", "carrot_yaw",
    "goal_x", "goal_y", "goal_yaw", "pose_x", "pose_y", "pose_yaw",
    "pose_yaw_objective", "control_interval", "delta_t",
)


# --------------------------------------------------------------------------------------
# parameters (srv.py:49-75 declares them with these defaults; README.md:53-84 is the sample)
# --------------------------------------------------------------------------------------
@dataclass
class MpcParams:
    acc_x_limit: float = 0.5
    acc_y_limit: float = 0.5
    acc_theta_limit: float = 0.5
    min_vel_x: float = -0.5
    min_vel_y: float = -0.5
    min_vel_trans: float = 0.5          # declared, never used (srv.py:55,84)
    min_vel_theta: float = -0.5
    max_vel_x: float = 0.5
    max_vel_y: float = 0.5
    max_vel_trans: float = 0.5
    max_vel_theta: float = 0.5
    w_trans: float = 0.5
    w_orient: float = 0.5
    w_control: float = 0.5
    w_terminal: float = 0.5
    w_costmap: float = 0.5
    w_footprint: float = 2000           # integer default in the reference (srv.py:68)
    waiting_time: float = 3.0           # never used as the threshold (srv.py:380 hard-codes 3.0)
    low_pass_gain: float = 0.5
    opt_tolerance: float = 1e-5
    prediction_horizon: float = 0.5
    control_steps: int = 3

    @classmethod
    def readme_sample(cls, **over):
        """The sample parameter file of the reference (README.md:53-84)."""
        p = cls(acc_x_limit=2.5, acc_y_limit=2.5, acc_theta_limit=3.0,
                min_vel_x=-0.7, min_vel_y=-0.7, min_vel_trans=-0.7, min_vel_theta=-0.7,
                max_vel_x=0.7, max_vel_y=0.7, max_vel_trans=0.7, max_vel_theta=0.7,
                w_trans=0.82, w_orient=0.50, w_control=0.05, w_terminal=0.05,
                w_footprint=0, w_costmap=0.05, waiting_time=3.0, low_pass_gain=0.5,
                opt_tolerance=1e-3, prediction_horizon=0.8, control_steps=3)
        for k, v in over.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p

    @property
    def dt(self):                       # srv.py:137
        return self.prediction_horizon / self.control_steps

    def as_dict(self):
        return asdict(self)


@dataclass
class Problem:
    """One Optimizer request, planar (yaw) form.  Mirrors ``neompc_request``."""
    vel_x: float = 0.0
    vel_y: float = 0.0
    vel_theta: float = 0.0
    carrot_x: float = 0.0
    carrot_y: float = 0.0
    carrot_yaw: float = 0.0
    goal_x: float = 0.0
    goal_y: float = 0.0
    goal_yaw: float = 0.0
    pose_x: float = 0.0
    pose_y: float = 0.0
    pose_yaw: float = 0.0
    pose_yaw_objective: float = 0.0
    control_interval: float = 1.0 / 30.0
    delta_t: float = 0.0
    instance_id: int = 0

    @classmethod
    def from_record(cls, rec):
        kw = {}
        for f in REQUEST_FIELDS:
            kw[f] = float(_get(rec, f))
        try:
            kw["instance_id"] = int(_get(rec, "instance_id"))
        except (KeyError, AttributeError, ValueError, IndexError):
            pass
        return cls(**kw)


def _get(rec, name):
    if isinstance(rec, dict):
        return rec[name]
    if hasattr(rec, "dtype") and getattr(rec.dtype, "names", None):
        return rec[name]
    return getattr(rec, name)


# --------------------------------------------------------------------------------------
# quaternion helpers  (srv.py:160-180, :182-196)
# --------------------------------------------------------------------------------------
def euler_yaw(x, y, z, w):
    """Yaw of ``euler_from_quaternion`` (srv.py:176-178)."""
    t3 = +2.0 * (w * z + x * y)
    t4 = +1.0 - 2.0 * (y * y + z * z)
    return math.atan2(t3, t4)


def quat_from_yaw(yaw):
    """``quaternion_from_euler(0, 0, yaw)`` (srv.py:182-196) returned as (x, y, z, w)."""
    cy = math.cos(yaw * 0.5)
    sy = math.sin(yaw * 0.5)
    # reference returns q = [w, x, y, z] with roll = pitch = 0
    return 0.0 * cy, 0.0 * cy, sy * 1.0 * 1.0, cy * 1.0 * 1.0


def quirk_yaw(pose_quat, goal_quat):
    """The yaw the objective's costmap rollout starts from: current-pose x, y, z with the
    GOAL pose's w (srv.py:213).  Quaternions are (x, y, z, w)."""
    return euler_yaw(pose_quat[0], pose_quat[1], pose_quat[2], goal_quat[3])


# --------------------------------------------------------------------------------------
# objective (srv.py:204-269) and constraint (srv.py:157-158): scalar, bit-exact restatement
# --------------------------------------------------------------------------------------
def objective(params: MpcParams, costmap, footprint_world, prob, cmd_vel, moving_footprint=None, bilinear=False):
    """J(cmd_vel).  ``footprint_world`` is the polygon the reference holds in
    ``self.footprint`` (world-frame vertices, list of (x, y)); because of the aliasing at
    srv.py:227/241-244 it never moves, so its cost is the same at every step.
    ``prob`` supplies the hoisted yaws: carrot_yaw (srv.py:211), goal_yaw (:212),
    pose_yaw_objective (:213, the goal-w quirk).

    ``moving_footprint`` (robot-frame vertices) selects the OPT-IN mode NEOMPC_FOOTPRINT_MOVING
     — NOT the reference's behaviour: the polygon is placed at each predicted
    pose (pos_x, pos_y, odom_yaw) of the costmap rollout (srv.py:234-236), which is what the loop at
    srv.py:238-244 sets out to do, and tested there with the same ``== 1.0`` rule (srv.py:262-263).

    ``bilinear=True`` selects the OPT-IN mode NEOMPC_COSTMAP_BILINEAR — NOT the reference's
    behaviour: the costmap term of srv.py:257-260 becomes (w_costmap c^2 + (1000 - w_costmap) l^2)/N with c and l the
    bilinear interpolations of the cell cost and of the lethal indicator; equal to the reference's term at cell centres."""
    n_steps = params.control_steps
    dt = params.dt
    cost_total = 0
    x = 0.0
    y = 0.0
    z = 0.0
    target_yaw = _get(prob, "carrot_yaw")
    final_yaw = _get(prob, "goal_yaw")
    odom_yaw = _get(prob, "pose_yaw_objective")
    v0x, v0y, v0z = _get(prob, "vel_x"), _get(prob, "vel_y"), _get(prob, "vel_theta")
    curr_pos = np.array((_get(prob, "carrot_x"), _get(prob, "carrot_y")))      # srv.py:219
    pos_x = _get(prob, "pose_x")
    pos_y = _get(prob, "pose_y")
    fp_cost = None

    for i in range(n_steps):                                                   # srv.py:224
        vx, vy, om = cmd_vel[0 + 3 * i], cmd_vel[1 + 3 * i], cmd_vel[2 + 3 * i]
        z += om * dt                                                           # :230
        x += (vx * np.cos(z) * dt - vy * np.sin(z) * dt)                       # :231
        y += (vx * np.sin(z) * dt + vy * np.cos(z) * dt)                       # :232
        odom_yaw += om * dt                                                    # :234
        pos_x += vx * np.cos(odom_yaw) * dt - vy * np.sin(odom_yaw) * dt       # :235
        pos_y += vx * np.sin(odom_yaw) * dt + vy * np.cos(odom_yaw) * dt       # :236

        mx1, my1 = costmap.getWorldToMap(pos_x, pos_y)                         # :246
        cell = costmap.getCost(mx1, my1)
        costmap_cost = 0 + cell ** 2                                           # :225,247

        step_dist_error = np.linalg.norm(curr_pos - np.array((x, y)))          # :250
        step_orient_error = target_yaw - z                                     # :251
        cost_total += ((params.w_trans * step_dist_error ** 2)
                       + (params.w_orient * step_orient_error ** 2)) / n_steps  # :252
        cost_total += params.w_control * (np.linalg.norm(
            np.array((v0x, v0y, v0z)) - np.array((vx, vy, om)))) / n_steps      # :253-254

        if bilinear:
            cb, lb = costmap.bilinear_at_world(pos_x, pos_y)[:2]
            cost_total += (params.w_costmap * float(cb) ** 2 + (1000 - params.w_costmap) * float(lb) ** 2) / n_steps
        elif cell == 1.0:                                                      # :257
            cost_total += costmap_cost * 1000 / n_steps                        # :258
        else:
            cost_total += params.w_costmap * costmap_cost / n_steps            # :260

        if moving_footprint is not None:
            fp_cost = costmap.getFootprintCost(_Poly(footprint_at(moving_footprint, pos_x, pos_y, odom_yaw)))
        elif fp_cost is None:
            fp_cost = costmap.getFootprintCost(_Poly(footprint_world))
        if fp_cost == 1.0:                                                     # :262
            cost_total += (fp_cost ** 2) * params.w_footprint / n_steps        # :263

    goal_xy = np.array((_get(prob, "goal_x"), _get(prob, "goal_y")))
    step_dist_error = np.linalg.norm(curr_pos - goal_xy)                       # :266
    step_orient_error = final_yaw - z                                          # :267
    cost_total += ((params.w_trans * step_dist_error ** 2)
                   + (params.w_orient * step_orient_error ** 2)) * params.w_terminal  # :268
    return cost_total


def f_constraint(params: MpcParams, u, index):
    """``max_vel_trans - sqrt(vx_i^2 + vy_i^2)`` (srv.py:157-158)."""
    return params.max_vel_trans - (np.sqrt((u[0 + index * 3]) * (u[0 + index * 3])
                                           + (u[1 + index * 3]) * (u[1 + index * 3])))


class _Pt:
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x, self.y = x, y


class _Poly:
    def __init__(self, pts):
        self.points = [p if hasattr(p, "x") else _Pt(p[0], p[1]) for p in (pts or [])]


def footprint_at(footprint_robot, x, y, yaw):
    """Robot-frame polygon placed at a predicted pose (moving-footprint mode; numpy float64 like the
    rest of the objective)."""
    c, s = np.cos(yaw), np.sin(yaw)
    return [(x + (fx * c - fy * s), y + (fx * s + fy * c)) for fx, fy in footprint_robot]


def footprint_world(footprint_robot, pose_x, pose_y, pose_yaw):
    """World-frame footprint polygon at the current pose — what nav2 publishes on
    ``/local_costmap/published_footprint`` and the reference stores at srv.py:154-155.
    (The transform itself is nav2's, not the reference's: declared, float64.)"""
    c, s = math.cos(pose_yaw), math.sin(pose_yaw)
    return [(pose_x + (fx * c - fy * s), pose_y + (fx * s + fy * c)) for fx, fy in footprint_robot]


# --------------------------------------------------------------------------------------
# the NLP solve (srv.py:125-134 bounds/constraints, :363-364 minimize call)
# --------------------------------------------------------------------------------------
def make_bounds_and_constraints(params: MpcParams):
    bnds, cons = [], []
    for i in range(params.control_steps):                                      # srv.py:130-134
        bnds.append((params.min_vel_x, params.max_vel_x))
        bnds.append((params.min_vel_y, params.max_vel_y))
        bnds.append((params.min_vel_theta, params.max_vel_theta))
        cons.append({"type": "ineq", "fun": partial(f_constraint, params, index=i)})
    return bnds, cons


def slsqp_solve(params: MpcParams, costmap, fp_world, prob, x0=None, ftol=None, maxiter=None, moving_footprint=None,
                bilinear=False):
    """Exactly the reference's call (srv.py:363-364): SLSQP, finite-difference gradients."""
    if x0 is None:
        x0 = np.zeros(params.control_steps * 3)                                # srv.py:136
    bnds, cons = make_bounds_and_constraints(params)
    opts = {"ftol": params.opt_tolerance if ftol is None else ftol, "disp": False}
    if maxiter is not None:
        opts["maxiter"] = maxiter
    fun = partial(objective, params, costmap, fp_world, prob, moving_footprint=moving_footprint, bilinear=bilinear)
    return minimize(fun, np.array(x0, dtype=np.float64), method="SLSQP",
                    bounds=bnds, constraints=cons, options=opts)


# --------------------------------------------------------------------------------------
# post-solve pieces of optimizer()
# --------------------------------------------------------------------------------------
def initial_guess_update(n_steps, init_guess, guess):
    """Shift the plan left by one step; tail = (already low-passed) first control (srv.py:198-202)."""
    for i in range(0, n_steps - 1):
        init_guess[0 + 3 * i:3 + 3 * i] = guess[3 + 3 * i:6 + 3 * i]
    init_guess[0 + 3 * (n_steps - 1):3 + 3 * (n_steps - 1)] = guess[0:3]
    return init_guess


def collision_check(params: MpcParams, costmap, fp_world, prob, x):
    """Re-roll the plan from the current pose with the TRUE current yaw (srv.py:312-347).
    Returns (hit, collision_footprint): ``hit`` -> the reference sets self.collision = True
    (latched); ``collision_footprint`` is assigned either way."""
    pos_x = _get(prob, "pose_x")
    pos_y = _get(prob, "pose_y")
    odom_yaw = _get(prob, "pose_yaw")                                          # :317
    dt = params.dt
    hit = False
    for i in range(params.control_steps):                                     # :323
        odom_yaw += x[2 + 3 * i] * dt
        pos_x += x[3 * i] * np.cos(odom_yaw) * dt - x[1 + 3 * i] * np.sin(odom_yaw) * dt
        pos_y += x[3 * i] * np.sin(odom_yaw) * dt + x[1 + 3 * i] * np.cos(odom_yaw) * dt
        mx1, my1 = costmap.getWorldToMap(pos_x, pos_y)
        col = costmap.getCost(mx1, my1)
        if col >= 0.99:                                                        # :338
            hit = True
            break
    fp_hit = costmap.getFootprintCost(_Poly(fp_world)) == 1.0                  # :343
    return hit, fp_hit


def local_plan(params: MpcParams, pose_x, pose_y, pose_yaw, x):
    """``publishLocalPlan(x)`` (srv.py:271-310): the poses of the published Path as rows
    (x, y, qz, qw) — N + 1 of them; row 0 is the start pose (position only, default orientation,
    srv.py:288-291).  (pose_x, pose_y, pose_yaw) stand for the TF lookup map -> base_link (:274-286)."""
    dt = params.dt
    pos_x, pos_y, yaw = pose_x, pose_y, pose_yaw
    rows = [(pos_x, pos_y, 0.0, 1.0)]
    for i in range(params.control_steps):                                     # :293
        yaw += x[2 + 3 * i] * dt                                               # :295
        pos_x += x[3 * i] * np.cos(yaw) * dt - x[1 + 3 * i] * np.sin(yaw) * dt  # :296
        pos_y += x[3 * i] * np.sin(yaw) * dt + x[1 + 3 * i] * np.cos(yaw) * dt  # :297
        q = quat_from_yaw(yaw)                                                 # :301 (x, y, z, w)
        rows.append((pos_x, pos_y, q[2], q[3]))
    return np.array(rows, dtype=np.float64)


class OracleServer:
    """The per-call state machine of ``MpcOptimizationServer.optimizer`` (srv.py:349-403) for ONE
    robot instance.  Wall-clock ``delta_t`` (srv.py:369-371) is an explicit request field."""

    def __init__(self, params: MpcParams, costmap, footprint_robot):
        self.params = params
        self.costmap = costmap
        self.footprint_robot = list(footprint_robot)
        n = params.control_steps
        self.initial_guess = np.zeros(n * 3)                                   # srv.py:136
        self.last_control = [0, 0, 0]                                          # :117
        self.waiting_time = params.waiting_time                                # :103
        self.collision = False                                                 # :148
        self.collision_footprint = False                                       # :149
        self.old_goal = None                                                   # :146 (never equals a real goal)
        self.last_result = None

    def tick(self, prob, solver=None):
        """One service call.  ``solver(x0, prob, fp_world) -> (x, success)`` can replace SLSQP
        (used to test the epilogue with the GPU's solution).  Returns (vx, vy, omega)."""
        p = self.params
        n = p.control_steps
        goal = (_get(prob, "goal_x"), _get(prob, "goal_y"), _get(prob, "goal_yaw"))
        new_goal = self.old_goal != goal
        if new_goal:                                                           # :358-361
            self.initial_guess = np.zeros(n * 3)
            self.last_control = [0, 0, 0]
            self.waiting_time = 0.0
        fp_world = footprint_world(self.footprint_robot, _get(prob, "pose_x"),
                                   _get(prob, "pose_y"), _get(prob, "pose_yaw"))
        if solver is None:
            res = slsqp_solve(p, self.costmap, fp_world, prob, self.initial_guess)   # :363-364
            xx, success = res.x, bool(res.success)
            self.last_result = res
        else:
            xx, success = solver(self.initial_guess.copy(), prob, fp_world)
            xx = np.array(xx, dtype=np.float64)
        self.solution = xx.copy()
        for i in range(0, 3):                                                  # :366-367
            xx[i] = xx[i] * p.low_pass_gain + self.last_control[i] * (1 - p.low_pass_gain)
        delta_t = _get(prob, "delta_t")                                        # :369-371
        hit, fp_hit = collision_check(p, self.costmap, fp_world, prob, xx)     # :372
        if hit:
            self.collision = True
        self.collision_footprint = fp_hit
        ci = _get(prob, "control_interval")
        if self.collision or self.collision_footprint:                         # :374-382
            out = [0.0, 0.0, 0.0]
            self.waiting_time += delta_t
            if self.waiting_time >= 3.0:
                self.collision = False
                self.waiting_time = 0.0
        else:                                                                  # :385-391
            acc = (p.acc_x_limit, p.acc_y_limit, p.acc_theta_limit)
            out = []
            for i in range(3):
                t = np.fmin(xx[i], self.last_control[i] + acc[i] * ci)
                out.append(np.fmax(t, self.last_control[i] - acc[i] * ci))
        self.last_control = [out[0], out[1], out[2]]                           # :393-395
        if success:                                                            # :397-400
            self.initial_guess = initial_guess_update(n, self.initial_guess, xx)
        else:
            self.initial_guess = xx
        self.old_goal = goal                                                   # :402
        self.new_goal = new_goal
        return float(out[0]), float(out[1]), float(out[2])


# --------------------------------------------------------------------------------------
# vectorised float64 versions (fast checks of whole batches; NOT bit-exact, ~1e-15 rel)
# --------------------------------------------------------------------------------------
def _col(reqs, name):
    return np.asarray(reqs[name], dtype=np.float64)


def rollout_batch(params: MpcParams, reqs, U, yaw_field="pose_yaw_objective"):
    """Base-frame (x, y, z) and world (px, py) after each step for a batch.  U: [B, 3N]."""
    n = params.control_steps
    dt = params.dt
    U = np.asarray(U, dtype=np.float64).reshape(len(U), n, 3)
    z = np.cumsum(U[:, :, 2] * dt, axis=1)
    c, s = np.cos(z), np.sin(z)
    x = np.cumsum((U[:, :, 0] * c - U[:, :, 1] * s) * dt, axis=1)
    y = np.cumsum((U[:, :, 0] * s + U[:, :, 1] * c) * dt, axis=1)
    oy = _col(reqs, yaw_field)[:, None] + z
    co, so = np.cos(oy), np.sin(oy)
    px = _col(reqs, "pose_x")[:, None] + np.cumsum((U[:, :, 0] * co - U[:, :, 1] * so) * dt, axis=1)
    py = _col(reqs, "pose_y")[:, None] + np.cumsum((U[:, :, 0] * so + U[:, :, 1] * co) * dt, axis=1)
    return x, y, z, px, py


def moving_footprint_lethal(params: MpcParams, costmap, reqs, U, footprint_robot):
    """bool[B, N]: footprint at the predicted pose of step i in collision (moving-footprint mode)."""
    n = params.control_steps
    U = np.asarray(U, dtype=np.float64).reshape(len(U), n, 3)
    _, _, z, px, py = rollout_batch(params, reqs, U)
    yaw = _col(reqs, "pose_yaw_objective")[:, None] + z
    out = np.zeros((len(U), n), dtype=bool)
    for b in range(len(U)):
        for i in range(n):
            out[b, i] = costmap.getFootprintCost(_Poly(footprint_at(footprint_robot, px[b, i], py[b, i],
                                                                   yaw[b, i]))) == 1.0
    return out


def objective_batch(params: MpcParams, costmap, reqs, U, fp_lethal=None, moving_footprint=None, bilinear=False):
    """Vectorised J for a batch (same formula as ``objective``).  ``fp_lethal``: bool[B], whether
    the current footprint cost == 1.0 (None -> all False).  ``costmap`` may be None (free space).
    ``moving_footprint``: robot-frame polygon -> the opt-in moving-footprint mode (then fp_lethal is ignored)."""
    p = params
    n = p.control_steps
    U = np.asarray(U, dtype=np.float64).reshape(len(U), n, 3)
    x, y, z, px, py = rollout_batch(p, reqs, U)
    cx, cy = _col(reqs, "carrot_x")[:, None], _col(reqs, "carrot_y")[:, None]
    d2 = (cx - x) ** 2 + (cy - y) ** 2
    oe = _col(reqs, "carrot_yaw")[:, None] - z
    J = ((p.w_trans * d2 + p.w_orient * oe ** 2) / n).sum(axis=1)
    v0 = np.stack([_col(reqs, "vel_x"), _col(reqs, "vel_y"), _col(reqs, "vel_theta")], axis=1)
    J += (p.w_control * np.sqrt(((v0[:, None, :] - U) ** 2).sum(axis=2)) / n).sum(axis=1)
    if costmap is not None and bilinear:
        c, l = costmap.bilinear_at_world(px, py)[:2]
        J += ((p.w_costmap * c ** 2 + (1000.0 - p.w_costmap) * l ** 2) / n).sum(axis=1)
    elif costmap is not None:
        c = costmap.cost_at_world(px, py)
        J += (np.where(c == 1.0, 1000.0, p.w_costmap) * c ** 2 / n).sum(axis=1)
    if moving_footprint is not None and costmap is not None:
        J += moving_footprint_lethal(p, costmap, reqs, U, moving_footprint).sum(axis=1) * (p.w_footprint / n)
    elif fp_lethal is not None:
        J += np.where(np.asarray(fp_lethal, dtype=bool), 1.0 * p.w_footprint, 0.0)
    dg2 = (_col(reqs, "carrot_x") - _col(reqs, "goal_x")) ** 2 + (_col(reqs, "carrot_y") - _col(reqs, "goal_y")) ** 2
    fe = _col(reqs, "goal_yaw") - z[:, -1]
    J += (p.w_trans * dg2 + p.w_orient * fe ** 2) * p.w_terminal
    return J


def gradient_batch(params: MpcParams, reqs, U, eps_control=0.0, bilinear_costmap=None):
    """Analytic gradient of the smooth part of J (costmap / footprint terms are piecewise
    constant -> 0 a.e.).  Control term: (u - v0)/sqrt(|u - v0|^2 + eps^2), 0 at the kink.
    ``bilinear_costmap``: a GridCostmap -> adds the gradient of the opt-in bilinear costmap term."""
    p = params
    n = p.control_steps
    dt = p.dt
    B = len(U)
    U = np.asarray(U, dtype=np.float64).reshape(B, n, 3)
    z = np.cumsum(U[:, :, 2] * dt, axis=1)
    c, s = np.cos(z), np.sin(z)
    dx = (U[:, :, 0] * c - U[:, :, 1] * s) * dt
    dy = (U[:, :, 0] * s + U[:, :, 1] * c) * dt
    x, y = np.cumsum(dx, axis=1), np.cumsum(dy, axis=1)
    cx, cy = _col(reqs, "carrot_x")[:, None], _col(reqs, "carrot_y")[:, None]
    gx = -2.0 * p.w_trans * (cx - x) / n
    gy = -2.0 * p.w_trans * (cy - y) / n
    gz = -2.0 * p.w_orient * (_col(reqs, "carrot_yaw")[:, None] - z) / n
    gz[:, -1] += -2.0 * p.w_orient * p.w_terminal * (_col(reqs, "goal_yaw") - z[:, -1])
    if bilinear_costmap is not None:
        # world position = pose + R(yaw0) (x, y): d/dx = cos(yaw0) d/dwx + sin(yaw0) d/dwy, d/dy = -sin d/dwx + cos d/dwy
        y0 = _col(reqs, "pose_yaw_objective")[:, None]
        px = _col(reqs, "pose_x")[:, None] + np.cos(y0) * x - np.sin(y0) * y
        py = _col(reqs, "pose_y")[:, None] + np.sin(y0) * x + np.cos(y0) * y
        cb, lb, dcx, dcy, dlx, dly = bilinear_costmap.bilinear_at_world(px, py)
        gwx = 2.0 * (p.w_costmap * cb * dcx + (1000.0 - p.w_costmap) * lb * dlx) / n
        gwy = 2.0 * (p.w_costmap * cb * dcy + (1000.0 - p.w_costmap) * lb * dly) / n
        gx = gx + np.cos(y0) * gwx + np.sin(y0) * gwy
        gy = gy - np.sin(y0) * gwx + np.cos(y0) * gwy
    Sx = np.cumsum(gx[:, ::-1], axis=1)[:, ::-1]
    Sy = np.cumsum(gy[:, ::-1], axis=1)[:, ::-1]
    Gz = gz - Sx * dy + Sy * dx
    SGz = np.cumsum(Gz[:, ::-1], axis=1)[:, ::-1]
    v0 = np.stack([_col(reqs, "vel_x"), _col(reqs, "vel_y"), _col(reqs, "vel_theta")], axis=1)
    diff = U - v0[:, None, :]
    r = np.sqrt((diff ** 2).sum(axis=2) + eps_control ** 2)
    with np.errstate(invalid="ignore", divide="ignore"):
        ctrl = np.where(r[:, :, None] > 0, p.w_control / n * diff / r[:, :, None], 0.0)
    G = np.empty_like(U)
    G[:, :, 0] = dt * (c * Sx + s * Sy) + ctrl[:, :, 0]
    G[:, :, 1] = dt * (-s * Sx + c * Sy) + ctrl[:, :, 1]
    G[:, :, 2] = dt * SGz + ctrl[:, :, 2]
    return G.reshape(B, 3 * n)

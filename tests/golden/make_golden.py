#!/usr/bin/env python3
"""Generate the golden vectors in this directory from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  It imports
``/root/reference/neo_mpc_planner2/mpc_optimization_server.py`` under the ROS stub modules of
``oracle/ros_stubs.py``, plugs the declared costmap fake (``oracle/costmap.py``) into it, drives the
reference's own ``objective`` / ``f_constraint`` / ``minimize`` call / ``optimizer`` handler and
writes inputs + outputs to JSON.  While generating, it asserts that the oracle restatement
(``oracle/mpc_oracle.py``) reproduces every number BIT-EXACTLY; ``tests/test_oracle_golden.py``
repeats that check against the committed files wherever the tests run.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import hashlib
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ros_stubs  # noqa: E402
import oracle  # noqa: E402
from oracle.costmap import GridCostmap, FreeSpaceCostmap  # noqa: E402
from oracle.mpc_oracle import footprint_world  # noqa: E402
from neo_mpc_planner2_b200 import workloads  # noqa: E402

README = oracle.MpcParams.readme_sample().as_dict()
FOOT = workloads.FOOTPRINT_RECT


class FakeClock:
    def __init__(self):
        self.now = 0.0

    def time(self):
        return self.now


def grid_random(seed):
    return np.random.default_rng(seed).integers(0, 101, (200, 200)).astype(np.uint8)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def set_request_state(mod, srv, prob):
    """Put a planar problem into the reference server's fields the way optimizer() does (srv.py:350-355)."""
    S = ros_stubs
    qc = oracle.quat_from_yaw(prob["carrot_yaw"])
    qg = oracle.quat_from_yaw(prob["goal_yaw"])
    qp = oracle.quat_from_yaw(prob["pose_yaw"])
    srv.carrot_pose = S.PoseStamped(pose=S.Pose(S.Point(prob["carrot_x"], prob["carrot_y"], 0.0), S.Quaternion(*qc)))
    srv.goal_pose = S.Pose(S.Point(prob["goal_x"], prob["goal_y"], 0.0), S.Quaternion(*qg))
    srv.current_pose = S.PoseStamped(pose=S.Pose(S.Point(prob["pose_x"], prob["pose_y"], 0.0), S.Quaternion(*qp)))
    srv.current_velocity = S.Twist(S.Vector3(prob["vel_x"], prob["vel_y"], 0.0), S.Vector3(0.0, 0.0, prob["vel_theta"]))
    return qp, qg


def make_request_msg(prob):
    S = ros_stubs
    r = S.OptimizerRequest()
    r.current_vel = S.Twist(S.Vector3(prob["vel_x"], prob["vel_y"], 0.0), S.Vector3(0.0, 0.0, prob["vel_theta"]))
    r.carrot_pose = S.PoseStamped(pose=S.Pose(S.Point(prob["carrot_x"], prob["carrot_y"], 0.0),
                                              S.Quaternion(*oracle.quat_from_yaw(prob["carrot_yaw"]))))
    r.goal_pose = S.Pose(S.Point(prob["goal_x"], prob["goal_y"], 0.0), S.Quaternion(*oracle.quat_from_yaw(prob["goal_yaw"])))
    r.current_pose = S.PoseStamped(pose=S.Pose(S.Point(prob["pose_x"], prob["pose_y"], 0.0),
                                               S.Quaternion(*oracle.quat_from_yaw(prob["pose_yaw"]))))
    r.switch_opt = False
    r.control_interval = prob["control_interval"]
    return r


def set_footprint(srv, fp_world_pts):
    S = ros_stubs
    srv.footprint = S.Polygon(points=[S.Point32(x, y, 0.0) for x, y in fp_world_pts])


def finish_problem(prob):
    """Fill the hoisted yaws the oracle consumes, computed with the reference's own formulae."""
    qp = oracle.quat_from_yaw(prob["pose_yaw"])
    qg = oracle.quat_from_yaw(prob["goal_yaw"])
    qc = oracle.quat_from_yaw(prob["carrot_yaw"])
    out = dict(prob)
    # what the reference's euler_from_quaternion returns for these quaternions
    out["carrot_yaw"] = oracle.euler_yaw(*qc)
    out["goal_yaw"] = oracle.euler_yaw(*qg)
    out["pose_yaw_objective"] = oracle.quirk_yaw(qp, qg)
    out["pose_yaw_true"] = oracle.euler_yaw(*qp)
    return out


def oracle_prob(fin):
    p = {k: fin[k] for k in oracle.mpc_oracle.REQUEST_FIELDS}
    p["pose_yaw"] = fin["pose_yaw_true"]
    return p


def rand_problem(rng, extent=3.5):
    b = rng.uniform(-math.pi, math.pi)
    return dict(
        vel_x=rng.uniform(-0.3, 0.3), vel_y=rng.uniform(-0.3, 0.3), vel_theta=rng.uniform(-0.3, 0.3),
        carrot_x=0.4 * math.cos(b), carrot_y=0.4 * math.sin(b), carrot_yaw=rng.uniform(-1, 1),
        goal_x=rng.uniform(-4, 4), goal_y=rng.uniform(-4, 4), goal_yaw=rng.uniform(-math.pi, math.pi),
        pose_x=rng.uniform(-extent, extent), pose_y=rng.uniform(-extent, extent),
        pose_yaw=rng.uniform(-math.pi, math.pi),
        control_interval=1.0 / 30.0, delta_t=1.0 / 30.0)


def main():
    mod = ros_stubs.load_reference()
    out = {"generator": "tests/golden/make_golden.py",
           "reference": "neobotix/neo_mpc_planner2 @ 752184836e (unmodified mpc_optimization_server.py under ROS stubs)",
           "versions": {"numpy": np.__version__, "scipy": __import__("scipy").__version__,
                        "python": sys.version.split()[0]}}

    # ------------------------------------------------------------------ 1. known answers
    srv = ros_stubs.make_server(mod, README)
    srv.costmap_ros = FreeSpaceCostmap()
    kat = dict(vel_x=0.0, vel_y=0.0, vel_theta=0.0, carrot_x=0.4, carrot_y=0.1, carrot_yaw=0.3,
               goal_x=3.0, goal_y=1.0, goal_yaw=0.5, pose_x=1.0, pose_y=2.0, pose_yaw=0.2,
               control_interval=1.0 / 30.0, delta_t=1000.0)
    fin = finish_problem(kat)
    set_request_state(mod, srv, kat)
    set_footprint(srv, footprint_world(FOOT, kat["pose_x"], kat["pose_y"], fin["pose_yaw_true"]))
    u_probe = [0.1, -0.2, 0.3, 0.4, 0.05, -0.1, -0.3, 0.2, 0.25]
    J0 = float(srv.objective(np.zeros(9)))
    J1 = float(srv.objective(np.array(u_probe)))
    assert abs(J0 - 0.5010200000000001) < 1e-15, J0
    assert abs(J1 - 0.4607352460331115) < 1e-15, J1
    p_or = oracle.MpcParams.readme_sample()
    fpw = footprint_world(FOOT, kat["pose_x"], kat["pose_y"], fin["pose_yaw_true"])
    assert oracle.objective(p_or, FreeSpaceCostmap(), fpw, oracle_prob(fin), np.zeros(9)) == J0
    assert oracle.objective(p_or, FreeSpaceCostmap(), fpw, oracle_prob(fin), np.array(u_probe)) == J1
    res = mod.minimize(srv.objective, np.zeros(9), method="SLSQP", bounds=srv.bnds, constraints=srv.cons,
                       options={"ftol": srv.opt_tolerance, "disp": False})
    res_o = oracle.slsqp_solve(p_or, FreeSpaceCostmap(), fpw, oracle_prob(fin))
    assert np.array_equal(res.x, res_o.x) and res.fun == res_o.fun and res.nit == res_o.nit
    tight = mod.minimize(srv.objective, np.zeros(9), method="SLSQP", bounds=srv.bnds, constraints=srv.cons,
                         options={"ftol": 1e-12, "maxiter": 1000, "disp": False})
    # first optimizer() call
    clock = FakeClock()
    mod.time = clock
    clock.now = kat["delta_t"]
    resp = srv.optimizer(make_request_msg(kat), ros_stubs.OptimizerResponse())
    tw = resp.output_vel.twist
    out["kat"] = {
        "params": README, "problem": fin, "footprint_robot": FOOT, "u_probe": u_probe,
        "J_zero": J0, "J_probe": J1,
        "slsqp": {"x": res.x.tolist(), "fun": float(res.fun), "nit": int(res.nit), "nfev": int(res.nfev),
                  "status": int(res.status), "success": bool(res.success)},
        "slsqp_tight": {"x": tight.x.tolist(), "fun": float(tight.fun), "nit": int(tight.nit),
                        "status": int(tight.status)},
        "first_tick": {"output": [float(tw.linear.x), float(tw.linear.y), float(tw.angular.z)],
                       "next_initial_guess": np.asarray(srv.initial_guess).tolist()},
    }
    print("KAT ok: J0=%r J1=%r slsqp fun=%r nit=%d nfev=%d" % (J0, J1, res.fun, res.nit, res.nfev))

    # ------------------------------------------------------------------ 2. objective / constraint cases
    cases = []
    rng = np.random.default_rng(20261017)
    for n_steps, horizon in ((3, 0.8), (10, 0.8), (20, 0.8), (3, 0.5), (7, 1.2)):
        for k in range(12):
            over = dict(README)
            over.update(control_steps=n_steps, prediction_horizon=horizon)
            if k % 3 == 1:
                over.update(w_footprint=7.0, w_costmap=0.3, w_control=0.2)
            if k % 3 == 2:      # code defaults of srv.py:49-75 (integer w_footprint = 2000)
                over = dict(oracle.MpcParams().as_dict())
                over.update(control_steps=n_steps, prediction_horizon=horizon)
            gseed = int(rng.integers(0, 2**31))
            cells = grid_random(gseed)
            if k % 4 == 3:
                cells[:, :] = 100                     # everything lethal: exercises the ==1.0 branches
            cm = GridCostmap(cells, 0.05, -5.0, -5.0)
            srv = ros_stubs.make_server(mod, over)
            srv.costmap_ros = cm
            prob = rand_problem(rng)
            if k == 5:
                prob.update(pose_x=4.9, pose_y=-4.95)  # rollout leaves the map -> out-of-bounds cells
            fin = finish_problem(prob)
            set_request_state(mod, srv, prob)
            fpw = footprint_world(FOOT, prob["pose_x"], prob["pose_y"], fin["pose_yaw_true"])
            set_footprint(srv, fpw)
            lo = np.tile([over["min_vel_x"], over["min_vel_y"], over["min_vel_theta"]], n_steps)
            hi = np.tile([over["max_vel_x"], over["max_vel_y"], over["max_vel_theta"]], n_steps)
            u = rng.uniform(lo, hi)
            if k == 0:
                u[:] = 0.0
            if k == 7:                               # sit exactly on the control-term kink
                u[0:3] = [prob["vel_x"], prob["vel_y"], prob["vel_theta"]]
            J = float(srv.objective(u.copy()))
            cons = [float(srv.f_constraint(u, i)) for i in range(n_steps)]
            fp_cost = float(cm.getFootprintCost(srv.footprint))
            p_or = oracle.MpcParams(**over)
            Jo = oracle.objective(p_or, cm, fpw, oracle_prob(fin), u.copy())
            assert Jo == J, (n_steps, k, Jo, J)
            assert [float(oracle.f_constraint(p_or, u, i)) for i in range(n_steps)] == cons
            cases.append({"params": over, "grid_seed": gseed, "grid_all_lethal": bool(k % 4 == 3),
                          "grid_sha256": sha(cells), "problem": fin, "footprint_world": fpw,
                          "u": u.tolist(), "J": J, "constraints": cons, "footprint_cost": fp_cost})
    out["objective_cases"] = cases
    print("objective cases:", len(cases), "all bit-exact")

    # ------------------------------------------------------------------ 3. SLSQP solves on a C2-like map
    wl = workloads.config("c2", batch=64)
    cm = GridCostmap(wl.cells, wl.resolution, wl.origin_x, wl.origin_y)
    solves = []
    for n_steps in (3, 10):
        for k in range(6):
            over = dict(README)
            over.update(control_steps=n_steps, w_footprint=2000 if k % 2 else 0)
            rec = wl.requests[k + (0 if n_steps == 3 else 6)]
            prob = {f: float(rec[f]) for f in oracle.mpc_oracle.REQUEST_FIELDS}
            fin = finish_problem(prob)
            srv = ros_stubs.make_server(mod, over)
            srv.costmap_ros = cm
            set_request_state(mod, srv, prob)
            fpw = footprint_world(FOOT, prob["pose_x"], prob["pose_y"], fin["pose_yaw_true"])
            set_footprint(srv, fpw)
            res = mod.minimize(srv.objective, np.zeros(3 * n_steps), method="SLSQP", bounds=srv.bnds,
                               constraints=srv.cons, options={"ftol": srv.opt_tolerance, "disp": False})
            p_or = oracle.MpcParams(**over)
            ro = oracle.slsqp_solve(p_or, cm, fpw, oracle_prob(fin))
            assert np.array_equal(res.x, ro.x) and res.fun == ro.fun and res.nit == ro.nit and res.nfev == ro.nfev
            solves.append({"params": over, "workload": "c2", "index": int(k + (0 if n_steps == 3 else 6)),
                           "problem": fin, "footprint_world": fpw,
                           "x": res.x.tolist(), "fun": float(res.fun), "nit": int(res.nit),
                           "nfev": int(res.nfev), "status": int(res.status), "success": bool(res.success)})
    out["slsqp_cases"] = solves
    out["c2_grid_sha256"] = sha(wl.cells)
    print("slsqp cases:", len(solves), "all bit-exact")

    # ------------------------------------------------------------------ 4. multi-tick sequences through optimizer()
    sequences = []

    def run_sequence(name, over, cells, origin, ticks):
        cm = GridCostmap(cells, 0.05, origin[0], origin[1]) if cells is not None else FreeSpaceCostmap()
        srv = ros_stubs.make_server(mod, over)
        srv.costmap_ros = cm
        clock = FakeClock()
        mod.time = clock
        p_or = oracle.MpcParams(**over)
        osrv = oracle.OracleServer(p_or, cm, FOOT)
        recs = []
        for t in ticks:
            # the reference derives delta_t from its wall clock (srv.py:369-371): record the
            # float64 difference it will actually see
            t = dict(t)
            new_now = clock.now + t["delta_t"]
            t["delta_t"] = new_now - clock.now
            fin = finish_problem(t)
            fpw = footprint_world(FOOT, t["pose_x"], t["pose_y"], fin["pose_yaw_true"])
            set_footprint(srv, fpw)
            clock.now = new_now
            resp = srv.optimizer(make_request_msg(t), ros_stubs.OptimizerResponse())
            tw = resp.output_vel.twist
            o_ref = [float(tw.linear.x), float(tw.linear.y), float(tw.angular.z)]
            o_or = list(osrv.tick(oracle_prob(fin)))
            assert o_ref == o_or, (name, o_ref, o_or)
            assert np.array_equal(np.asarray(srv.initial_guess, dtype=float), osrv.initial_guess)
            assert bool(srv.collision) == osrv.collision and bool(srv.collision_footprint) == osrv.collision_footprint
            assert float(srv.waiting_time) == float(osrv.waiting_time)
            recs.append({"problem": fin, "output": o_ref,
                         "initial_guess_after": np.asarray(srv.initial_guess, dtype=float).tolist(),
                         "collision": bool(srv.collision), "collision_footprint": bool(srv.collision_footprint),
                         "waiting_time": float(srv.waiting_time),
                         "last_control": [float(v) for v in srv.last_control]})
        sequences.append({"name": name, "params": over, "footprint_robot": FOOT,
                          "grid": None if cells is None else {"sha256": sha(cells)},
                          "origin": list(origin), "ticks": recs})
        print("sequence %-28s %d ticks bit-exact" % (name, len(recs)))

    def drive(start, goal_of_tick, n_ticks, carrot=(0.4, 0.1, 0.3), dt_first=1000.0, dt=1.0 / 30.0,
              outputs_from=None):
        """Ticks with a pose that integrates the previous outputs (computed by the reference
        on the fly) are built lazily: here we just precompute a kinematically plausible drift."""
        ticks = []
        x, y, yaw = start
        for k in range(n_ticks):
            g = goal_of_tick(k)
            ticks.append(dict(vel_x=0.02 * k, vel_y=0.0, vel_theta=0.01 * k,
                              carrot_x=carrot[0], carrot_y=carrot[1], carrot_yaw=carrot[2],
                              goal_x=g[0], goal_y=g[1], goal_yaw=g[2],
                              pose_x=x, pose_y=y, pose_yaw=yaw,
                              control_interval=1.0 / 30.0, delta_t=dt_first if k == 0 else dt))
            x += 0.01 * math.cos(yaw)
            y += 0.01 * math.sin(yaw)
            yaw += 0.002
        return ticks

    # (a) free space, N=3, one goal change at tick 4 (reset of guess / last_control)
    over = dict(README)
    run_sequence("free_goal_change_n3", over, None, (0.0, 0.0),
                 drive((1.0, 2.0, 0.2), lambda k: (3.0, 1.0, 0.5) if k < 4 else (-2.0, 0.5, -1.0), 8))
    # (b) N=10 on the C2 map with the footprint weight on
    over = dict(README)
    over.update(control_steps=10, w_footprint=2000)
    p0 = wl.requests[20]
    run_sequence("c2map_n10", over, wl.cells, (wl.origin_x, wl.origin_y),
                 drive((float(p0["pose_x"]), float(p0["pose_y"]), float(p0["pose_yaw"])),
                       lambda k: (2.0, -1.0, 0.7), 5))
    # (c) wall ahead: predicted cell >= 0.99 -> stop, wait 3 s (delta_t = 0.8 s per tick), release
    cells = np.zeros((200, 200), dtype=np.uint8)
    cells[:, 112:] = 99          # inscribed band from x = 0.6 m
    cells[:, 116:] = 100
    over = dict(README)
    run_sequence("wall_ahead_stop_wait_release", over, cells, (-5.0, -5.0),
                 drive((0.3, 0.0, 0.0), lambda k: (4.0, 0.0, 0.0), 13, carrot=(0.4, 0.0, 0.0), dt=0.8))
    # (d) footprint already lethal (robot overlaps an obstacle): collision_footprint every tick
    cells = np.zeros((200, 200), dtype=np.uint8)
    cells[100:104, 106:110] = 100     # block at x in [0.3,0.5), y in [0,0.2) touches the footprint edge x=0.4
    over = dict(README)
    over.update(w_footprint=2000)
    run_sequence("footprint_lethal", over, cells, (-5.0, -5.0),
                 drive((0.0, 0.0, 0.0), lambda k: (4.0, 0.0, 0.0), 4, carrot=(-0.4, 0.0, 0.0)))
    # (e) code-default parameters (srv.py:49-75), tight tolerance 1e-5
    over = dict(oracle.MpcParams().as_dict())
    run_sequence("code_defaults_n3", over, None, (0.0, 0.0),
                 drive((0.0, 0.0, 1.0), lambda k: (1.0, 1.0, 0.0), 4, carrot=(0.3, -0.2, -0.4)))
    out["tick_sequences"] = sequences
    out["wall_grid"] = "cells[:,112:]=99; cells[:,116:]=100 on 200x200, res 0.05, origin (-5,-5)"
    out["footprint_grid"] = "cells[100:104,106:110]=100 on 200x200, res 0.05, origin (-5,-5)"

    path = os.path.join(HERE, "reference_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

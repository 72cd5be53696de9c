#!/usr/bin/env python3
"""Which dynamic parameters actually change the reference's solve?  Golden facts from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  ``cb_params`` (mpc_optimization_server.py:405-439) accepts 14
names, but the bounds list is built once at start-up (srv.py:125-133) and ``w_costmap`` / ``w_footprint`` are copied to
``*_scale`` attributes (srv.py:96-97) that the callback never touches.  This script loads the reference under the ROS
stand-ins, changes one parameter at a time through the reference's own callback and records whether the objective value,
the constraint value and the optimizer's result change.  ``neo_mpc_planner2_b200/ros_node.py`` mirrors exactly that
(``EFFECTIVE_IN_REFERENCE``); ``tests/test_ros_node.py`` checks the mirror against the file written here.

    python tests/golden/make_golden_dynamic_params.py
"""
from __future__ import annotations

import json
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ros_stubs  # noqa: E402
import oracle  # noqa: E402
from oracle.costmap import GridCostmap  # noqa: E402
from make_golden import set_request_state, set_footprint, make_request_msg, FakeClock  # noqa: E402

README = oracle.MpcParams.readme_sample().as_dict()
NAMES = ("min_vel_x", "min_vel_y", "min_vel_trans", "min_vel_theta", "max_vel_x", "max_vel_y", "max_vel_trans",
         "max_vel_theta", "w_trans", "w_orient", "w_control", "w_terminal", "w_costmap", "w_footprint")
NEW = dict(min_vel_x=-0.2, min_vel_y=-0.2, min_vel_trans=-0.2, min_vel_theta=-0.2, max_vel_x=0.2, max_vel_y=0.2,
           max_vel_trans=0.3, max_vel_theta=0.2, w_trans=1.7, w_orient=1.1, w_control=0.4, w_terminal=0.6,
           w_costmap=3.0, w_footprint=55.0)
PROB = dict(vel_x=0.1, vel_y=-0.05, vel_theta=0.2, carrot_x=0.4, carrot_y=0.1, carrot_yaw=0.3, goal_x=3.0, goal_y=1.0,
            goal_yaw=0.5, pose_x=0.5, pose_y=0.5, pose_yaw=0.2, control_interval=1.0 / 30.0, delta_t=1.0 / 30.0)
# scenario B starts from rest and has a whole second of acceleration headroom, so the accel clamp (srv.py:385-391)
# does not hide differences between solutions (for scenario A's input this scipy version returns the zero start unchanged)
PROB_B = dict(PROB, vel_x=0.0, vel_y=0.0, vel_theta=0.0, control_interval=1.0)


def fresh(mod, footprint_in_collision=True):
    """Scenario A (objective facts): the footprint polygon lies over lethal cells, so w_footprint would matter if it
    were effective.  Scenario B (optimizer() results): a free footprint and cells below 0.6, so no collision stop
    (srv.py:374-377) masks the solve."""
    params = dict(README, control_steps=3, w_footprint=7.0)
    srv = ros_stubs.make_server(mod, params)
    cells = np.random.default_rng(0).integers(0, 60, (200, 200)).astype(np.uint8)
    if footprint_in_collision:
        cells[110:112, 110:125] = 100
    srv.costmap_ros = GridCostmap(cells, 0.05, -5.0, -5.0)
    set_request_state(mod, srv, PROB)
    set_footprint(srv, [(0.9, 0.8), (0.1, 0.8), (0.1, 0.2), (0.9, 0.2)])            # world-frame polygon
    return srv


def main():
    mod = ros_stubs.load_reference()
    mod.time = FakeClock()
    P = ros_stubs.Parameter
    u = np.array([0.1, -0.2, 0.3, 0.4, 0.05, -0.1, -0.3, 0.2, 0.25])
    facts = {}
    for name in NAMES:
        srv = fresh(mod)
        assert srv.costmap_ros.getFootprintCost(srv.footprint) == 1.0
        j0, c0 = float(srv.objective(u)), float(srv.f_constraint(u, 0))
        srv = fresh(mod)
        change = [SimpleNamespace(name=name, value=NEW[name], type_=P.Type.DOUBLE)]
        assert srv.cb_params(change).successful
        j1, c1 = float(srv.objective(u)), float(srv.f_constraint(u, 0))
        srv = fresh(mod, footprint_in_collision=False)
        r0 = srv.optimizer(make_request_msg(PROB_B), ros_stubs.OptimizerResponse())
        x0 = (r0.output_vel.twist.linear.x, r0.output_vel.twist.linear.y, r0.output_vel.twist.angular.z)
        assert x0 != (0.0, 0.0, 0.0)
        srv = fresh(mod, footprint_in_collision=False)
        assert srv.cb_params(change).successful
        r1 = srv.optimizer(make_request_msg(PROB_B), ros_stubs.OptimizerResponse())
        x1 = (r1.output_vel.twist.linear.x, r1.output_vel.twist.linear.y, r1.output_vel.twist.angular.z)
        facts[name] = dict(new_value=NEW[name], objective_before=j0, objective_after=j1, objective_changed=j0 != j1,
                           constraint_changed=c0 != c1, result_before=x0, result_after=x1, result_changed=x0 != x1)
        print(f"{name:15s} objective {'changes' if j0 != j1 else 'same   '} constraint {'changes' if c0 != c1 else 'same   '}"
              f" result {'changes' if x0 != x1 else 'same'}")
    # a non-DOUBLE parameter is ignored altogether (srv.py:407)
    srv = fresh(mod)
    j0 = float(srv.objective(u))
    srv.cb_params([SimpleNamespace(name="w_trans", value=9.0, type_=2)])
    facts["_non_double_ignored"] = float(srv.objective(u)) == j0
    out = dict(generator="tests/golden/make_golden_dynamic_params.py",
               reference="neo_mpc_planner2/mpc_optimization_server.py:405-439 (cb_params), unmodified", facts=facts)
    with open(os.path.join(HERE, "dynamic_params_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("effective:", [n for n in NAMES if facts[n]["objective_changed"] or facts[n]["constraint_changed"]
                         or facts[n]["result_changed"]])


if __name__ == "__main__":
    main()

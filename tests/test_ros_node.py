"""The ROS 2 node shim for un-modified clients, exercised under the stand-in ROS modules of
oracle/ros_stubs.py (this image has no rclpy): same node name, service, topics and parameter names as the
reference; answers equal the library's direct answers; the published Path equals the oracle's publishLocalPlan."""
import os
import sys
import types

import numpy as np
import pytest

import oracle
from neo_mpc_planner2_b200 import workloads

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture()
def ros(monkeypatch):
    from oracle import ros_stubs
    saved = {k: v for k, v in sys.modules.items()}
    ros_stubs.install()
    sys.modules.pop("neo_mpc_planner2_b200.ros_node", None)
    yield ros_stubs
    for k in list(sys.modules):
        if k not in saved:
            del sys.modules[k]


class Recorder:
    def __init__(self):
        self.msgs = []

    def publish(self, m):
        self.msgs.append(m)


def make_request(S, r):
    rq = S.OptimizerRequest()
    rq.current_vel.linear.x, rq.current_vel.linear.y, rq.current_vel.angular.z = (float(r["vel_x"]), float(r["vel_y"]),
                                                                                   float(r["vel_theta"]))
    for pose, px, py, yaw in ((rq.carrot_pose.pose, "carrot_x", "carrot_y", "carrot_yaw"),
                              (rq.goal_pose, "goal_x", "goal_y", "goal_yaw"),
                              (rq.current_pose.pose, "pose_x", "pose_y", "pose_yaw")):
        pose.position.x, pose.position.y = float(r[px]), float(r[py])
        q = oracle.quat_from_yaw(float(r[yaw]))
        pose.orientation.x, pose.orientation.y, pose.orientation.z, pose.orientation.w = q
    rq.control_interval = float(r["control_interval"])
    return rq


def test_node_module_mirrors_reference_interface(ros):
    """Importable under the ROS stand-ins; parameter names and defaults are the reference's (srv.py:49-75)."""
    from neo_mpc_planner2_b200 import ros_node
    assert ros_node.REFERENCE_PARAMETERS == oracle.MpcParams().as_dict()
    assert set(ros_node.EFFECTIVE_IN_REFERENCE) <= set(ros_node.DYNAMIC_NAMES) <= set(ros_node.REFERENCE_PARAMETERS)
    assert issubclass(ros_node.MpcOptimizationServer, ros.Node)


def test_dynamic_parameter_semantics_match_the_reference(ros):
    """tests/golden/dynamic_params_golden.json records, from the UNMODIFIED reference, which of the 14 names its
    cb_params accepts (srv.py:405-439) actually change the objective, the constraint or the optimizer's result: the
    bounds are built once (srv.py:125-133) and w_costmap / w_footprint are copied at start-up (srv.py:96-97).  The node
    shim must treat exactly that set as effective in its reference-faithful mode."""
    import json
    from neo_mpc_planner2_b200 import ros_node
    with open(os.path.join(HERE, "golden", "dynamic_params_golden.json")) as f:
        facts = json.load(f)["facts"]
    names = [k for k in facts if not k.startswith("_")]
    assert set(names) == set(ros_node.DYNAMIC_NAMES)
    effective = {k for k in names if facts[k]["objective_changed"] or facts[k]["constraint_changed"] or facts[k]["result_changed"]}
    assert effective == set(ros_node.EFFECTIVE_IN_REFERENCE)
    for k in names:                     # no parameter changes the result without changing objective or constraint
        assert facts[k]["result_changed"] == (facts[k]["objective_changed"] or facts[k]["constraint_changed"])
    assert facts["_non_double_ignored"] is True


@pytest.mark.gpu
def test_node_answers_like_the_library(ros):
    from neo_mpc_planner2_b200 import ros_node
    from neo_mpc_planner2_b200.solver import BatchSolver
    from oracle.mpc_oracle import footprint_world
    wl = workloads.config("c2", batch=64)
    created = {}
    orig_service = ros.Node.create_service
    ros.Node.create_service = lambda self, typ, name, cb: created.setdefault("service", (typ, name, cb))
    subs = {}
    ros.Node.create_subscription = lambda self, typ, topic, cb, depth: subs.setdefault(topic, cb)
    try:
        ros.PARAM_OVERRIDES.update(wl.params)
        node = ros_node.MpcOptimizationServer()
        ros.PARAM_OVERRIDES.clear()
    finally:
        ros.Node.create_service = orig_service
    assert created["service"][0] is ros.Optimizer and created["service"][1] == "optimizer"
    assert set(subs) == {"/local_costmap/published_footprint", "/local_costmap/costmap", "/local_costmap/costmap_updates"}
    node.PubRaysPath = Recorder()
    # before a costmap and a footprint have arrived the node stands still instead of solving in free space
    early = created["service"][2](make_request(ros, wl.requests[0]), ros.OptimizerResponse())
    assert (early.output_vel.twist.linear.x, early.output_vel.twist.linear.y, early.output_vel.twist.angular.z) == (0.0, 0.0, 0.0)
    assert node.PubRaysPath.msgs == []
    # costmap and footprint arrive on their topics: first an EMPTY full grid, then the obstacles as update patches (what
    # nav2 sends while the costmap origin stands still, always_send_full_costmap: false)
    grid = types.SimpleNamespace(
        info=types.SimpleNamespace(width=wl.cells.shape[1], height=wl.cells.shape[0], resolution=wl.resolution,
                                   origin=types.SimpleNamespace(position=types.SimpleNamespace(x=wl.origin_x, y=wl.origin_y))),
        data=np.zeros(wl.cells.size, np.int8).tolist())
    subs["/local_costmap/costmap"](grid)
    H, W = wl.cells.shape
    for (y0, x0) in ((0, 0), (0, W // 2), (H // 2, 0), (H // 2, W // 2)):
        patch = wl.cells[y0:y0 + H // 2, x0:x0 + W // 2].astype(np.int8)
        subs["/local_costmap/costmap_updates"](ros.OccupancyGridUpdate(x=x0, y=y0, width=patch.shape[1], height=patch.shape[0],
                                                                       data=patch.ravel().tolist()))
    assert node.costmap_generation == 5 and np.array_equal(node._grid.view(np.uint8), wl.cells)
    p = oracle.MpcParams(**wl.params)
    with BatchSolver(wl.params) as direct:
        direct.load_workload(wl)
        direct.reserve_instances(1)
        for k in (0, 5, 9):
            r = wl.requests[k]
            fpw = footprint_world(wl.footprint, float(r["pose_x"]), float(r["pose_y"]), float(r["pose_yaw"]))
            subs["/local_costmap/published_footprint"](
                ros.PolygonStamped(polygon=ros.Polygon(points=[ros.Point32(x, y, 0.0) for x, y in fpw])))
            resp = created["service"][2](make_request(ros, r), ros.OptimizerResponse())
            one = wl.requests[k:k + 1].copy()
            one["instance_id"] = 0
            one["delta_t"] = 1e9                      # first call / new goal each time
            direct.reset_state()
            want, plan = direct.solve(one, want_plan=True)
            got = (resp.output_vel.twist.linear.x, resp.output_vel.twist.linear.y, resp.output_vel.twist.angular.z)
            assert np.abs(np.array(got) - np.array([want["vx"][0], want["vy"][0], want["omega"][0]])).max() <= 2e-4
            path = node.PubRaysPath.msgs[-1]
            assert path.header.frame_id == "map" and len(path.poses) == p.control_steps + 1
            mine = oracle.local_plan(p, float(r["pose_x"]), float(r["pose_y"]),
                                     oracle.euler_yaw(*oracle.quat_from_yaw(float(r["pose_yaw"]))), node.solution.astype(np.float64))
            got_path = np.array([[ps.pose.position.x, ps.pose.position.y, ps.pose.orientation.z, ps.pose.orientation.w]
                                 for ps in path.poses])
            assert np.abs(got_path - mine).max() <= 1e-5
    # dynamic parameters: the reference's callback semantics (srv.py:405-439)
    P = ros.Parameter
    mk = lambda n, v, t=P.Type.DOUBLE: types.SimpleNamespace(name=n, value=v, type_=t)
    res = node.cb_params([mk("w_trans", 1.5), mk("max_vel_x", 0.2), mk("w_costmap", 9.0), mk("w_orient", 0.9, 2)])
    assert res.successful
    assert node.params["w_trans"] == 1.5
    assert node.params["max_vel_x"] == wl.params["max_vel_x"]      # bounds are built once in the reference (srv.py:125-133)
    assert node.params["w_costmap"] == wl.params["w_costmap"]      # copied to w_costmap_scale at start-up (srv.py:96)
    assert node.params["w_orient"] == wl.params["w_orient"]        # not a DOUBLE: ignored (srv.py:407)
    assert float(node._solver.params["w_trans"]) == 1.5
    node.destroy_node()

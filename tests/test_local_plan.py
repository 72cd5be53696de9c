"""Predicted-path output: ``publishLocalPlan`` (srv.py:271-310).

Golden vectors come from the unmodified reference (tests/golden/make_golden_local_plan.py).  CPU: the oracle restatement
equals them bit for bit.  GPU: ``neompc_local_plan`` (float64 on the device) on the float32-rounded inputs the C ABI
carries agrees with the oracle on the same rounded inputs to 1e-7 (the only differences: prediction_horizon is a
float32 parameter, and CUDA's / glibc's cos and sin differ in the last bit), and with the float64 golden poses to 1e-5
(float32 rounding of a pose up to 40 m from the origin)."""
import json
import os

import numpy as np
import pytest

import oracle
from neo_mpc_planner2_b200.abi import REQUEST_DTYPE, STATELESS

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def cases():
    with open(os.path.join(HERE, "golden", "local_plan_golden.json")) as f:
        return json.load(f)["cases"]


def test_oracle_equals_reference_golden(cases):
    assert len(cases) >= 20
    for c in cases:
        p = oracle.MpcParams(**c["params"])
        got = oracle.local_plan(p, c["pose"][0], c["pose"][1], c["pose"][2], np.array(c["x"]))
        want = np.array(c["poses"])
        assert got.shape == (p.control_steps + 1, 4)
        assert (got[:, 0] == want[:, 0]).all() and (got[:, 1] == want[:, 1]).all()
        assert (want[:, 2] == 0.0).all() and (want[:, 3] == 0.0).all()
        assert (got[:, 2] == want[:, 4]).all() and (got[:, 3] == want[:, 5]).all()
        # unit quaternions about z; the first pose keeps the default orientation
        assert np.abs(want[:, 4] ** 2 + want[:, 5] ** 2 - 1.0).max() < 1e-15
        assert want[0, 4] == 0.0 and want[0, 5] == 1.0


@pytest.mark.gpu
def test_gpu_local_plan_golden(cases):
    from neo_mpc_planner2_b200.solver import BatchSolver
    by_params = {}
    for c in cases:
        by_params.setdefault(json.dumps(c["params"], sort_keys=True), []).append(c)
    for key, group in by_params.items():
        params = json.loads(key)
        p = oracle.MpcParams(**params)
        req = np.zeros(len(group), REQUEST_DTYPE)
        req["instance_id"] = STATELESS
        X = np.zeros((len(group), 3 * p.control_steps), np.float32)
        for k, c in enumerate(group):
            req["pose_x"][k], req["pose_y"][k], req["pose_yaw"][k] = c["pose"]
            X[k] = c["x"]
        with BatchSolver(params) as s:
            poses = s.local_plan(req, X)
        assert poses.shape == (len(group), p.control_steps + 1)
        for k, c in enumerate(group):
            mine = oracle.local_plan(p, float(req["pose_x"][k]), float(req["pose_y"][k]), float(req["pose_yaw"][k]),
                                     X[k].astype(np.float64))
            got = np.stack([poses[k]["x"], poses[k]["y"], poses[k]["qz"], poses[k]["qw"]], 1)
            assert np.abs(got - mine).max() <= 1e-7
            want = np.array(c["poses"])[:, [0, 1, 4, 5]]
            assert np.abs(got - want).max() <= 1e-5


@pytest.mark.gpu
def test_gpu_local_plan_of_a_solved_batch_and_server_mirror():
    from neo_mpc_planner2_b200 import workloads
    from neo_mpc_planner2_b200.solver import BatchSolver
    from neo_mpc_planner2_b200 import server as srv
    wl = workloads.config("c3", batch=512)
    p = oracle.MpcParams(**wl.params)
    with BatchSolver(wl.params) as s:
        s.load_workload(wl)
        out, plan = s.solve(wl.requests, want_plan=True)
        poses = s.local_plan(wl.requests, plan)
    for k in (0, 17, 511):
        r = wl.requests[k]
        mine = oracle.local_plan(p, float(r["pose_x"]), float(r["pose_y"]), float(r["pose_yaw"]), plan[k].astype(np.float64))
        got = np.stack([poses[k]["x"], poses[k]["y"], poses[k]["qz"], poses[k]["qw"]], 1)
        assert np.abs(got - mine).max() <= 1e-7
    # empty batch is a no-op
    with BatchSolver(wl.params) as s:
        assert s.local_plan(wl.requests[:0], plan[:0]).shape == (0, p.control_steps + 1)
    # the server mirror publishes the same path through its PubRaysPath (srv.py:107-108, :365)
    class Rec:
        last = None
        def publish(self, m):
            self.last = m
    server = srv.MpcOptimizationServer(wl.params)
    server.set_costmap(wl.cells, wl.resolution, wl.origin_x, wl.origin_y)
    server.footprint_callback(wl.footprint)
    server.PubRaysPath = Rec()
    r = wl.requests[3]
    rq = srv.OptimizerRequest()
    rq.current_vel.linear.x, rq.current_vel.linear.y, rq.current_vel.angular.z = float(r["vel_x"]), float(r["vel_y"]), float(r["vel_theta"])
    rq.carrot_pose.pose.position.x, rq.carrot_pose.pose.position.y = float(r["carrot_x"]), float(r["carrot_y"])
    rq.carrot_pose.pose.orientation = srv.quaternion_from_yaw(float(r["carrot_yaw"]))
    rq.goal_pose.position.x, rq.goal_pose.position.y = float(r["goal_x"]), float(r["goal_y"])
    rq.goal_pose.orientation = srv.quaternion_from_yaw(float(r["goal_yaw"]))
    rq.current_pose.pose.position.x, rq.current_pose.pose.position.y = float(r["pose_x"]), float(r["pose_y"])
    rq.current_pose.pose.orientation = srv.quaternion_from_yaw(float(r["pose_yaw"]))
    rq.control_interval = float(r["control_interval"])
    server.optimizer(rq)
    path = server.PubRaysPath.last
    assert path is server.local_plan and path.frame_id == "map" and len(path.poses) == p.control_steps + 1
    mine = oracle.local_plan(p, float(r["pose_x"]), float(r["pose_y"]),
                             oracle.euler_yaw(0.0, 0.0, rq.current_pose.pose.orientation.z, rq.current_pose.pose.orientation.w),
                             server.solution.astype(np.float64))
    got = np.array([[ps.pose.position.x, ps.pose.position.y, ps.pose.orientation.z, ps.pose.orientation.w] for ps in path.poses])
    assert np.abs(got - mine).max() <= 1e-5
    server.close()

"""The C++ controller plugin (neo_mpc_planner::NeoMpcPlanner over libneompc): builds against the header stand-ins,
exports its factory, refuses to run without a GPU, and on a GPU returns the same twists as the Python mirror of the
reference's server fed with the same requests."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "neo_mpc_planner2_b200", "plugin")


@pytest.fixture(scope="module")
def built():
    from neo_mpc_planner2_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    subprocess.check_call(["make", "-C", PLUGIN], stdout=subprocess.DEVNULL)
    return PLUGIN


def test_plugin_builds_and_exports_factory(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(built, "libneo_mpc_planner2.so")], text=True)
    assert "neompc_plugin_create" in out
    for method in ("computeVelocityCommands", "configure", "setPlan", "setSpeedLimit", "cleanup", "activate", "deactivate"):
        assert method in out, method


def test_plugin_has_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = subprocess.run([os.path.join(built, "plugin_demo")], capture_output=True, text=True)
    assert res.returncode == 2 and "ControllerException" in res.stderr and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_plugin_matches_python_server(built):
    res = subprocess.run([os.path.join(built, "plugin_demo")], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    ticks = json.loads(res.stdout.strip().splitlines()[-1])
    assert len(ticks) == 5

    from neo_mpc_planner2_b200.abi import README_SAMPLE
    from neo_mpc_planner2_b200 import server as S
    srv = S.MpcOptimizationServer(dict(README_SAMPLE), device=0)
    srv.set_costmap(np.zeros((200, 200), np.uint8), 0.05, -5.0, -5.0, encoding=1)
    srv.footprint_callback([(0.4, 0.3), (-0.4, 0.3), (-0.4, -0.3), (0.4, -0.3)])
    plan = [(0.1 * i, 0.02 * i) for i in range(41)]
    px = py = 0.0
    vel = (0.0, 0.0, 0.0)
    for k in range(5):
        d = [math.hypot(x - px, y - py) for x, y in plan]
        start = int(np.argmin(d))
        pick = next((i for i in range(start, len(plan)) if d[i] >= 0.4), len(plan) - 1)
        req = S.OptimizerRequest()
        req.current_vel.linear.x, req.current_vel.linear.y, req.current_vel.angular.z = vel
        req.carrot_pose.pose.position.x = plan[pick][0] - px          # robot yaw is 0 in the demo
        req.carrot_pose.pose.position.y = plan[pick][1] - py
        req.carrot_pose.pose.orientation = S.quaternion_from_yaw(0.2)
        req.goal_pose.position.x, req.goal_pose.position.y = plan[-1]
        req.goal_pose.orientation = S.quaternion_from_yaw(0.2)
        req.current_pose.pose.position.x, req.current_pose.pose.position.y = px, py
        req.control_interval = 1.0 / 30.0
        out = srv.optimizer(req).output_vel.twist
        got = (out.linear.x, out.linear.y, out.angular.z)
        assert max(abs(a - b) for a, b in zip(got, ticks[k])) <= 1e-5, (k, got, ticks[k])
        vel = got
        px += got[0] / 30.0
        py += got[1] / 30.0
    srv.close()

"""The C++ controller plugin (neo_mpc_planner::NeoMpcPlanner over libneompc): builds against the header stand-ins,
exports its factory, refuses to run without a GPU, throws the reference's ControllerExceptions, and on a GPU its closed
loop is replayed tick by tick against the ORACLE: carrot bookkeeping and the Optimizer request against
oracle.carrot_oracle.select_carrot (cpp:66-246), the twist against oracle.OracleServer — the reference's optimizer() state
machine (srv.py:349-403) — fed with the plan the plugin's solver returned."""
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


def scene(shift=0):
    """The demo's costmap (plugin_demo.cpp: paint_scene), nav2 raw costs, integer arithmetic."""
    W = H = 200
    bx0, bx1, by0, by1 = 92 + shift, 100 + shift, 112, 120
    x = np.arange(W)[None, :]
    y = np.arange(H)[:, None]
    dx = np.where(x < bx0, bx0 - x, np.where(x >= bx1, x - bx1 + 1, 0))
    dy = np.where(y < by0, by0 - y, np.where(y >= by1, y - by1 + 1, 0))
    d = np.maximum(dx, dy)
    v = np.where(d == 0, 254, np.where(d <= 2, 253, 252 - 18 * (d - 2)))
    return np.maximum(v, 0).astype(np.uint8)


def demo_plan(n=60):
    t = 0.05 * np.arange(n + 1)
    yaw = np.arctan2(0.3 * t, 1.0) + np.where(np.arange(n + 1) > 30, 1.2, 0.0)
    # the plugin extracts the yaw from the pose quaternion (z = sin(yaw/2), w = cos(yaw/2)) with atan2
    qz, qw = np.sin(0.5 * yaw), np.cos(0.5 * yaw)
    yaw_q = np.arctan2(2.0 * qw * qz, 1.0 - 2.0 * qz * qz)
    return np.stack([-1.0 + t, 0.15 * t * t, yaw_q], axis=1)


@pytest.mark.gpu
def test_plugin_closed_loop_against_the_oracle(built):
    import oracle
    from oracle.carrot_oracle import select_carrot, footprint_raw_cost
    from oracle.costmap import GridCostmap, ENC_NAV2_RAW
    from oracle.mpc_oracle import REQUEST_FIELDS
    from neo_mpc_planner2_b200.abi import README_SAMPLE
    res = subprocess.run([os.path.join(built, "plugin_demo")], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    run = json.loads(res.stdout.strip().splitlines()[-1])
    ticks = run["ticks"]
    assert len(ticks) == 12
    fp = [(0.4, 0.3), (-0.4, 0.3), (-0.4, -0.3), (0.4, -0.3)]
    plan = demo_plan()
    params = dict(README_SAMPLE)
    p = oracle.MpcParams(**params)
    cms = {s: GridCostmap(scene(s), 0.05, -5.0, -5.0, ENC_NAV2_RAW) for s in (0, 2)}
    srv = oracle.OracleServer(p, cms[0], fp)
    plan_start, slow_down = 0, True                                   # h:162: slow_down_ starts true
    saw_pruning = saw_slow = False
    for k, t in enumerate(ticks):
        cm = cms[0] if k <= 8 else cms[2]                             # the demo moves the obstacle after tick 7's command
        srv.costmap = cm
        # the costmap is uploaded on the first tick and again right after it changed — not on the others
        assert t["uploaded"] == (1 if k in (0, 8) else 0), (k, t["uploaded"])
        x, y, yaw = t["pose"]
        fc = footprint_raw_cost(cm, fp, x, y, yaw)
        o = select_carrot(plan, plan_start, (x, y, yaw), slow_down, 0.3, 0.45, 0.25, 200 * 0.05 / 2.0, fc)
        status, begin, carrot_index, flags = t["info"]
        assert status == o["status"] == 0
        assert begin == o["begin"] and carrot_index == o["carrot_index"], (k, t["info"], o)
        assert bool(flags & 1) == o["closer_to_goal"] and bool(flags & 2) == o["slow_down"]
        assert (flags >> 8) & 0xFF == fc
        rq = t["request"]
        assert abs(rq["carrot_x"] - o["carrot"][0]) <= 1e-6 and abs(rq["carrot_y"] - o["carrot"][1]) <= 1e-6
        assert abs(rq["carrot_yaw"] - o["carrot"][2]) <= 1e-6
        assert abs(rq["goal_x"] - plan[-1, 0]) <= 1e-6 and abs(rq["goal_y"] - plan[-1, 1]) <= 1e-6
        assert abs(rq["control_interval"] - 1.0 / 30.0) <= 1e-8 and rq["instance_id"] == 0
        saw_pruning |= begin > 0
        saw_slow |= o["slow_down"]
        plan_start, slow_down = o["begin"], o["slow_down"]
        # the reference's optimizer() on the plugin's request, driven with the plan the plugin's solver returned
        prob = oracle.Problem(**{f: rq[f] for f in REQUEST_FIELDS})
        ok = t["response"]["status"] != 1
        want = srv.tick(prob, solver=lambda x0, pr, fpw: (np.array(t["plan"], dtype=np.float64), ok))
        assert max(abs(a - b) for a, b in zip(want, t["twist"])) <= 1e-6, (k, want, t["twist"])
        assert bool(t["response"]["flags"] & 1) == srv.collision
    assert saw_pruning, "the closed loop never moved past the first plan pose"
    # the predicted path of the last tick (publishLocalPlan, srv.py:271-310)
    last = ticks[-1]
    lp = oracle.local_plan(p, last["request"]["pose_x"], last["request"]["pose_y"], last["request"]["pose_yaw"],
                           np.array(last["plan"], dtype=np.float64))
    assert np.abs(np.array(run["local_plan"]) - lp).max() <= 1e-6


@pytest.mark.gpu
def test_plugin_refuses_a_plan_in_another_frame(built):
    res = subprocess.run([os.path.join(built, "plugin_demo"), "frames"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2 and "Unable to transform robot pose into global plan's frame" in res.stderr


@pytest.mark.gpu
def test_plugin_full_tick_latency(built):
    """Full computeVelocityCommands ticks (checksum of the costmap, front half, solve, one synchronise): a 60x60 local
    costmap and a 1000x1000 one, unchanged and changing every tick."""
    for w, h, bound_us in ((60, 60, 400.0), (1000, 1000, 3000.0)):
        res = subprocess.run([os.path.join(built, "plugin_demo"), "latency", str(w), str(h)], capture_output=True, text=True,
                             timeout=300)
        assert res.returncode == 0, res.stderr
        d = json.loads(res.stdout.strip().splitlines()[-1])
        print("\nplugin tick latency", d)
        assert d["costmap_unchanged_us"]["median"] <= bound_us
        assert d["costmap_unchanged_us"]["median"] <= d["costmap_new_every_tick_us"]["median"]

"""SURVEY §8f row N2 — batched carrot selection / request construction (reference src/NeoMpcPlanner.cpp:66-246)
on the device vs its numpy restatement (oracle/carrot_oracle.py; unpinned: the C++ plugin needs ROS 2 to build)."""
import math

import numpy as np
import pytest

from oracle.carrot_oracle import select_carrot, footprint_raw_cost, raw_byte_table, STATUS_OK
from oracle.costmap import GridCostmap, ENC_NAV2_RAW, ENC_OCCUPANCY
from neo_mpc_planner2_b200 import workloads
from neo_mpc_planner2_b200.abi import TICK_DTYPE, README_SAMPLE, STATELESS


def _plan(n=900):
    s = np.linspace(0.0, 1.0, n)
    x = -20.0 + 40.0 * s
    y = 8.0 * np.sin(2.5 * np.pi * s)
    yaw = np.arctan2(np.gradient(y), np.gradient(x))
    return np.stack([x, y, yaw], 1)


def _ticks(plan, n, seed):
    rng = np.random.default_rng(seed)
    k = rng.integers(0, len(plan), n)
    t = np.zeros(n, TICK_DTYPE)
    t["pose_x"] = plan[k, 0] + rng.uniform(-0.4, 0.4, n)
    t["pose_y"] = plan[k, 1] + rng.uniform(-0.4, 0.4, n)
    t["pose_yaw"] = plan[k, 2] + rng.uniform(-1.6, 1.6, n)
    t["vel_x"], t["vel_y"], t["vel_theta"] = rng.uniform(-0.3, 0.3, (3, n))
    t["plan_start"] = np.maximum(0, k - rng.integers(0, 60, n))
    t["plan_start"][::7] = np.minimum(len(plan) - 1, k[::7] + 5)      # already pruned past the closest pose
    t["slow_down"] = rng.integers(0, 2, n)
    t["delta_t"] = 1.0 / 30.0
    return t


def test_oracle_properties():
    plan = _plan(300)
    for seed in range(3):
        t = _ticks(plan, 50, seed)
        for r in t:
            o = select_carrot(plan, int(r["plan_start"]), (r["pose_x"], r["pose_y"], r["pose_yaw"]), bool(r["slow_down"]),
                              0.3, 0.5, 0.4, 25.0, 0)
            assert o["status"] == STATUS_OK and o["begin"] >= r["plan_start"] and o["carrot_index"] >= o["begin"]
            hyp = math.hypot(o["carrot"][0], o["carrot"][1])
            la = 0.4 if o["closer_to_goal"] else (0.5 if not r["slow_down"] else 0.3)
            assert hyp >= la - 1e-12 or o["carrot_index"] == len(plan) - 1
            if o["carrot_index"] > o["begin"]:       # the pose before the carrot was still too close
                px, py = plan[o["carrot_index"] - 1, :2]
                assert math.hypot(px - r["pose_x"], py - r["pose_y"]) < la + 1e-12
    assert raw_byte_table(ENC_OCCUPANCY)[[0, 1, 98, 99, 100, 255]].tolist() == [0, 1, 252, 253, 254, 255]


@pytest.mark.gpu
@pytest.mark.parametrize("encoding", [ENC_OCCUPANCY, ENC_NAV2_RAW])
def test_build_requests_matches_oracle(encoding):
    from neo_mpc_planner2_b200.solver import BatchSolver
    wl = workloads.config("c3", batch=64)
    cells = wl.cells.copy()
    if encoding == ENC_NAV2_RAW:
        cells = raw_byte_table(ENC_OCCUPANCY)[cells].astype(np.uint8)
        cells[400:420, 380:400] = 255                                  # a patch of unknown space
    cm = GridCostmap(cells, wl.resolution, wl.origin_x, wl.origin_y, encoding)
    plan = _plan()
    ticks = _ticks(plan, 3000, 5)
    la = (np.float32(0.3), np.float32(0.5), np.float32(0.4))
    with BatchSolver(dict(README_SAMPLE, control_steps=10)) as s:
        s.set_costmap(cells, wl.resolution, wl.origin_x, wl.origin_y, encoding)
        s.set_footprint(wl.footprint)
        s.set_plan(plan)
        cp = s.carrot_params(*[float(v) for v in la], controller_frequency=30.0)
        s.reserve_instances(100 + len(ticks))                          # ids 100 .. 100 + n - 1 need their state rows
        reqs, info = s.build_requests(ticks, cp, first_instance_id=100)
        out = s.solve(np.where(True, reqs, reqs))                      # the requests feed the solver as they are
    mtd = max(cells.shape) * wl.resolution / 2.0
    near = 0
    for i, r in enumerate(ticks):
        fc = footprint_raw_cost(cm, wl.footprint, r["pose_x"], r["pose_y"], r["pose_yaw"])
        o = select_carrot(plan, int(r["plan_start"]), (r["pose_x"], r["pose_y"], r["pose_yaw"]), bool(r["slow_down"]),
                          float(la[0]), float(la[1]), float(la[2]), mtd, fc)
        got = info[i]
        same = (got["status"] == o["status"] and got["plan_start"] == o["begin"] and
                got["carrot_index"] == o["carrot_index"] and
                (got["flags"] & 1) == int(o["closer_to_goal"]) and ((got["flags"] >> 1) & 1) == int(o["slow_down"]))
        if not same:
            near += 1                                                  # only float64 last-bit ties may differ
            continue
        assert (got["flags"] >> 8) & 0xFF == fc, (i, (got["flags"] >> 8) & 0xFF, fc)
        q = reqs[i]
        assert abs(q["carrot_x"] - o["carrot"][0]) <= 1e-6 and abs(q["carrot_y"] - o["carrot"][1]) <= 1e-6
        assert abs(q["carrot_yaw"] - o["carrot"][2]) <= 1e-6
        assert q["goal_x"] == np.float32(plan[-1, 0]) and q["goal_yaw"] == np.float32(plan[-1, 2])
        assert q["pose_x"] == np.float32(r["pose_x"]) and q["pose_yaw"] == np.float32(r["pose_yaw"])
        assert q["vel_theta"] == r["vel_theta"] and q["instance_id"] == 100 + i
        want_q = workloads.quirk_yaw_planar(float(r["pose_yaw"]), float(plan[-1, 2]))
        assert abs(q["pose_yaw_objective"] - want_q) <= 1e-6
        assert q["control_interval"] == np.float32(1.0 / 30.0)
    assert near <= 3, near
    assert ((info["flags"] >> 1) & 1).sum() > 0 and (info["flags"] & 1).sum() >= 0
    assert np.isfinite(out["vx"]).all()


@pytest.mark.gpu
def test_pruning_state_over_ticks():
    """Feeding plan_start / slow_down back tick after tick reproduces the reference's erase-as-you-go pruning."""
    from neo_mpc_planner2_b200.solver import BatchSolver
    plan = _plan(400)
    n = 64
    rng = np.random.default_rng(9)
    ticks = np.zeros(n, TICK_DTYPE)
    k = rng.integers(0, 100, n)
    ticks["pose_x"], ticks["pose_y"], ticks["pose_yaw"] = plan[k, 0], plan[k, 1] + 0.1, plan[k, 2]
    ticks["slow_down"] = 1                                             # slow_down_ starts true (h:162)
    state = [dict(start=0, slow=True) for _ in range(n)]
    with BatchSolver(dict(README_SAMPLE)) as s:
        s.set_plan(plan)
        cp = s.carrot_params(0.3, 0.5, 0.4, 30.0)
        for step in range(6):
            reqs, info = s.build_requests(ticks, cp, first_instance_id=STATELESS)
            for i in range(n):
                o = select_carrot(plan, state[i]["start"], (ticks["pose_x"][i], ticks["pose_y"][i], ticks["pose_yaw"][i]),
                                  state[i]["slow"], float(np.float32(0.3)), 0.5, float(np.float32(0.4)), 1e300, 0)
                assert info["plan_start"][i] == o["begin"] and info["carrot_index"][i] == o["carrot_index"]
                assert ((info["flags"][i] >> 1) & 1) == int(o["slow_down"])
                state[i] = dict(start=o["begin"], slow=o["slow_down"])
            # move every robot along its carrot and feed the state back
            c, sn = np.cos(ticks["pose_yaw"]), np.sin(ticks["pose_yaw"])
            ticks["pose_x"] += 0.5 * (c * reqs["carrot_x"] - sn * reqs["carrot_y"])
            ticks["pose_y"] += 0.5 * (sn * reqs["carrot_x"] + c * reqs["carrot_y"])
            ticks["plan_start"] = info["plan_start"]
            ticks["slow_down"] = (info["flags"] >> 1) & 1
        assert (ticks["plan_start"] > k).all()                         # the plans were pruned as the robots advanced

#!/usr/bin/env python3
"""Golden vectors for the predicted-path output from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  Loads
``/root/reference/neo_mpc_planner2/mpc_optimization_server.py`` under the ROS stubs of ``oracle/ros_stubs.py``, replaces the
server's TF buffer by one that answers (the stub's default raises, which makes ``publishLocalPlan`` return early) and
its publisher by one that records, calls the reference's own ``publishLocalPlan(x)`` (srv.py:271-310) and writes
inputs + the published poses to ``local_plan_golden.json``.  Asserts ``==`` against ``oracle.local_plan`` while doing so.

    python tests/golden/make_golden_local_plan.py
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

README = oracle.MpcParams.readme_sample().as_dict()


class AnsweringBuffer:
    def __init__(self):
        self.pose = (0.0, 0.0, 0.0)

    def lookup_transform(self, target, source, when):
        assert (target, source) == ("map", "base_link")
        x, y, yaw = self.pose
        q = oracle.quat_from_yaw(yaw)
        return SimpleNamespace(transform=SimpleNamespace(
            translation=SimpleNamespace(x=x, y=y, z=0.0),
            rotation=SimpleNamespace(x=q[0], y=q[1], z=q[2], w=q[3])))


class Recorder:
    def __init__(self):
        self.last = None

    def publish(self, msg):
        self.last = msg


def main():
    mod = ros_stubs.load_reference()
    rng = np.random.default_rng(77)
    cases = []
    for n_steps, horizon in ((3, 0.8), (10, 0.8), (20, 0.8), (7, 0.5)):
        params = dict(README, control_steps=n_steps, prediction_horizon=horizon)
        srv = ros_stubs.make_server(mod, params)
        srv.tf_buffer = AnsweringBuffer()
        srv.PubRaysPath = Recorder()
        p = oracle.MpcParams(**params)
        for _ in range(6):
            pose = (float(rng.uniform(-40, 40)), float(rng.uniform(-40, 40)), float(rng.uniform(-np.pi, np.pi)))
            x = rng.uniform(-0.7, 0.7, 3 * n_steps)
            srv.tf_buffer.pose = pose
            srv.publishLocalPlan(x)
            path = srv.PubRaysPath.last
            assert path.header.frame_id == "map" and len(path.poses) == n_steps + 1
            rows = [[ps.pose.position.x, ps.pose.position.y, ps.pose.orientation.x, ps.pose.orientation.y,
                     ps.pose.orientation.z, ps.pose.orientation.w] for ps in path.poses]
            # the yaw the reference extracts from the TF quaternion (srv.py:286) is what the oracle is given
            yaw0 = oracle.euler_yaw(*oracle.quat_from_yaw(pose[2]))
            mine = oracle.local_plan(p, pose[0], pose[1], yaw0, x)
            for r, m in zip(rows, mine):
                assert r[0] == m[0] and r[1] == m[1] and r[2] == 0.0 and r[3] == 0.0 and r[4] == m[2] and r[5] == m[3], (r, m)
            cases.append(dict(params=params, pose=[pose[0], pose[1], yaw0], x=x.tolist(), poses=rows))
    out = dict(generator="tests/golden/make_golden_local_plan.py",
               reference="neo_mpc_planner2/mpc_optimization_server.py:271-310 (publishLocalPlan), unmodified",
               layout="poses: [x, y, qx, qy, qz, qw] per pose, N + 1 poses", cases=cases)
    with open(os.path.join(HERE, "local_plan_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(cases), "cases; oracle.local_plan == reference on every number")


if __name__ == "__main__":
    main()

"""Drives the UNMODIFIED reference (oracle/_ref, see build_ref.py) for the bench's CPU arms: one solve is the reference's
own ``minimize(self.objective, self.initial_guess, method='SLSQP', bounds=self.bnds, constraints=self.cons,
options={'ftol': self.opt_tolerance, 'disp': False})`` (srv.py:363-364) on a ``MpcOptimizationServer`` object built under
the ROS stand-in modules, with the request fields set the way ``optimizer()`` sets them (srv.py:350-355), the footprint the
way ``footprint_callback`` does (srv.py:154-155) and the declared costmap fake plugged in where ``Costmap2d(self)`` goes
(srv.py:118).  TEST / BENCH INFRASTRUCTURE ONLY — never imported by the product."""
from __future__ import annotations

import os

import numpy as np

from . import ros_stubs
from .mpc_oracle import quat_from_yaw, footprint_world


def available() -> bool:
    return os.path.exists(ros_stubs.REF_PYC)


class ReferenceSolver:
    def __init__(self, params: dict, costmap, footprint_robot):
        self.mod = ros_stubs.load_reference_compiled()
        self.srv = ros_stubs.make_server(self.mod, dict(params))
        self.srv.costmap_ros = costmap
        self.footprint_robot = list(footprint_robot)

    def load(self, prob):
        """Request -> server fields (srv.py:350-355); footprint topic (srv.py:154-155)."""
        S, srv = ros_stubs, self.srv
        qc, qg, qp = quat_from_yaw(prob.carrot_yaw), quat_from_yaw(prob.goal_yaw), quat_from_yaw(prob.pose_yaw)
        srv.carrot_pose = S.PoseStamped(pose=S.Pose(S.Point(prob.carrot_x, prob.carrot_y, 0.0), S.Quaternion(*qc)))
        srv.goal_pose = S.Pose(S.Point(prob.goal_x, prob.goal_y, 0.0), S.Quaternion(*qg))
        srv.current_pose = S.PoseStamped(pose=S.Pose(S.Point(prob.pose_x, prob.pose_y, 0.0), S.Quaternion(*qp)))
        srv.current_velocity = S.Twist(S.Vector3(prob.vel_x, prob.vel_y, 0.0), S.Vector3(0.0, 0.0, prob.vel_theta))
        fpw = footprint_world(self.footprint_robot, prob.pose_x, prob.pose_y, prob.pose_yaw)
        srv.footprint = S.Polygon(points=[S.Point32(x, y, 0.0) for x, y in fpw])

    def solve(self, prob):
        """Cold start, exactly the reference's call (srv.py:363-364)."""
        srv = self.srv
        self.load(prob)
        srv.initial_guess = np.zeros(srv.no_ctrl_steps * 3)                   # srv.py:136 / :359
        return self.mod.minimize(srv.objective, srv.initial_guess, method="SLSQP", bounds=srv.bnds,
                                 constraints=srv.cons, options={"ftol": srv.opt_tolerance, "disp": False})

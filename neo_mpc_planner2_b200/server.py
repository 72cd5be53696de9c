"""Host-side mirror of the reference's service interface for the hot path.

``MpcOptimizationServer.optimizer(request, response)`` has the reference's name, argument meaning and
behaviour (mpc_optimization_server.py:349-403): it takes a ``neo_srvs2/srv/Optimizer`` request (any object with
the same attribute tree: ``current_vel``, ``carrot_pose``, ``goal_pose``, ``current_pose``, ``switch_opt``,
``control_interval``) and fills ``response.output_vel.twist``.  The work is done by libneompc on the GPU
(``neompc_solve_msgs`` with n = 1); the per-robot state the reference keeps in ``self`` lives in the library's
device-side instance row.  There is no scipy and no CPU path here.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np

from .abi import MSG_DTYPE, REQUEST_DTYPE, README_SAMPLE, ENC_OCCUPANCY
from .solver import BatchSolver


# ---- minimal message types (same attribute trees as geometry_msgs / neo_srvs2; used when rclpy is absent)
@dataclass
class Vector3:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0


@dataclass
class Quaternion:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0
    w: float = 1.0


@dataclass
class Pose:
    position: Vector3 = field(default_factory=Vector3)
    orientation: Quaternion = field(default_factory=Quaternion)


@dataclass
class PoseStamped:
    pose: Pose = field(default_factory=Pose)


@dataclass
class Twist:
    linear: Vector3 = field(default_factory=Vector3)
    angular: Vector3 = field(default_factory=Vector3)


@dataclass
class TwistStamped:
    twist: Twist = field(default_factory=Twist)


@dataclass
class OptimizerRequest:
    current_vel: Twist = field(default_factory=Twist)
    carrot_pose: PoseStamped = field(default_factory=PoseStamped)
    goal_pose: Pose = field(default_factory=Pose)
    current_pose: PoseStamped = field(default_factory=PoseStamped)
    switch_opt: bool = False
    control_interval: float = 0.0


@dataclass
class OptimizerResponse:
    output_vel: TwistStamped = field(default_factory=TwistStamped)


@dataclass
class Path:
    frame_id: str = "map"                       # srv.py:309
    poses: list = field(default_factory=list)   # PoseStamped


def _pose7(pose):
    p, q = pose.position, pose.orientation
    return [p.x, p.y, p.z, q.x, q.y, q.z, q.w]


def request_to_msg(request, delta_t, instance_id=0):
    """Marshal an Optimizer request (attribute tree of cpp:240-246) into a ``neompc_optimizer_request`` record."""
    m = np.zeros(1, MSG_DTYPE)
    v = request.current_vel
    m["current_vel"][0] = [v.linear.x, v.linear.y, v.linear.z, v.angular.x, v.angular.y, v.angular.z]
    m["carrot_pose"][0] = _pose7(request.carrot_pose.pose)
    m["goal_pose"][0] = _pose7(request.goal_pose)
    m["current_pose"][0] = _pose7(request.current_pose.pose)
    m["control_interval"] = request.control_interval
    m["delta_t"] = delta_t
    m["instance_id"] = instance_id
    m["switch_opt"] = 1 if request.switch_opt else 0
    return m


def requests_to_msgs(reqs):
    """Planar float32 request records -> float64 quaternion messages (z = sin(yaw/2), w = cos(yaw/2))."""
    reqs = np.asarray(reqs, dtype=REQUEST_DTYPE)
    m = np.zeros(len(reqs), MSG_DTYPE)
    m["current_vel"][:, 0] = reqs["vel_x"]
    m["current_vel"][:, 1] = reqs["vel_y"]
    m["current_vel"][:, 5] = reqs["vel_theta"]
    for name, px, py, yaw in (("carrot_pose", "carrot_x", "carrot_y", "carrot_yaw"),
                              ("goal_pose", "goal_x", "goal_y", "goal_yaw"),
                              ("current_pose", "pose_x", "pose_y", "pose_yaw")):
        a = reqs[yaw].astype(np.float64)
        m[name][:, 0] = reqs[px]
        m[name][:, 1] = reqs[py]
        m[name][:, 5] = np.sin(a * 0.5)
        m[name][:, 6] = np.cos(a * 0.5)
    m["control_interval"] = reqs["control_interval"]
    m["delta_t"] = reqs["delta_t"]
    m["instance_id"] = reqs["instance_id"]
    return m


class MpcOptimizationServer:
    """Drop-in for the reference's ``MpcOptimizationServer`` as far as the hot path goes: construct with the
    same parameter names (srv.py:49-75), feed costmap and footprint, call ``optimizer(request, response)``."""

    def __init__(self, params=None, device: int = 0, **over):
        p = dict(README_SAMPLE if params is None else params)
        p.update(over)
        self.params = p
        self._solver = BatchSolver(p, device=device)
        self._solver.reserve_instances(1)
        self.last_time = 0.0                      # srv.py:138
        self.last_response = None
        self.solution = None                      # x.x of srv.py:363: the solver's controls, 3*control_steps
        self.local_plan = Path()                  # srv.py:109: what publishLocalPlan publishes on /mpc_local_plan
        self.PubRaysPath = None                   # srv.py:107-108: any object with .publish(path); None = keep only

    # -- environment inputs (the reference gets these from ROS topics)
    def set_costmap(self, cells, resolution, origin_x, origin_y, encoding=ENC_OCCUPANCY):
        self._solver.set_costmap(cells, resolution, origin_x, origin_y, encoding)

    def footprint_callback(self, robot_frame_xy):
        """The reference stores the world-frame polygon from /local_costmap/published_footprint (srv.py:154-155);
        here the robot-frame polygon is given once and placed at each request's current pose on the device."""
        self._solver.set_footprint(robot_frame_xy)

    def cb_params(self, **changes):
        """Dynamic parameter update (srv.py:405-439)."""
        self.params.update(changes)
        self._solver.set_params(**changes)

    # -- the service handler
    def optimizer(self, request, response=None):
        if response is None:
            response = OptimizerResponse()
        current_time = time.time()                # srv.py:369-371 (the first call sees a huge delta_t, as there)
        delta_t = current_time - self.last_time
        self.last_time = current_time
        msg = request_to_msg(request, delta_t, instance_id=0)
        out, plan = self._solver.solve_msgs(msg, want_plan=True)
        response.output_vel.twist.linear.x = float(out["vx"][0])     # srv.py:375-377 / :389-391
        response.output_vel.twist.linear.y = float(out["vy"][0])
        response.output_vel.twist.angular.z = float(out["omega"][0])
        self.last_response = out[0]
        self.solution = plan[0]
        self.publishLocalPlan(plan[0], request.current_pose.pose)                # srv.py:365
        return response

    def publishLocalPlan(self, x, start_pose=None):
        """srv.py:271-310: roll the plan x from the robot pose and publish it as a Path.  The reference reads the pose
        from TF (map -> base_link, srv.py:274-286); the mirror takes the request's current_pose.  The rollout itself
        runs on the device (neompc_local_plan)."""
        if start_pose is None:
            return
        q = start_pose.orientation
        req = np.zeros(1, REQUEST_DTYPE)
        req["pose_x"], req["pose_y"] = start_pose.position.x, start_pose.position.y
        req["pose_yaw"] = math.atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z))   # srv.py:176-178
        poses = self._solver.local_plan(req, np.asarray(x, np.float32)[None, :])[0]
        self.local_plan = Path(poses=[
            PoseStamped(pose=Pose(Vector3(float(p["x"]), float(p["y"]), 0.0),
                                  Quaternion(0.0, 0.0, float(p["qz"]), float(p["qw"])))) for p in poses])
        if self.PubRaysPath is not None:
            self.PubRaysPath.publish(self.local_plan)

    def close(self):
        self._solver.close()


def quaternion_from_yaw(yaw):
    """quaternion_from_euler(0, 0, yaw) (srv.py:182-196) as a Quaternion message."""
    return Quaternion(0.0, 0.0, math.sin(yaw * 0.5), math.cos(yaw * 0.5))

"""``mpc_optimization_server`` as a ROS 2 node for UN-MODIFIED clients.

The reference's C++ plugin calls the service ``optimizer`` (type ``neo_srvs2/srv/Optimizer``, client created at
``src/NeoMpcPlanner.cpp:308``, called ``:248-250``).  This node registers the same service under the same node name with
the same parameters (``mpc_optimization_server.py:49-75``) and answers it from libneompc on the GPU, so an existing Nav2
stack keeps working without touching the plugin: only the Python executable is swapped.  What it mirrors:

  * parameters, declared with the reference's names and defaults                    srv.py:49-75, read :78-103
  * service 'optimizer'                                                              srv.py:105, handler :349-403
  * publisher 'local_plan' (nav_msgs/Path, the predicted path)                       srv.py:107, :271-310
  * subscription '/local_costmap/published_footprint' (geometry_msgs/PolygonStamped) srv.py:140-144, :154-155
  * the costmap: the reference reads it through neo_nav2_py_costmap2D's Costmap2d(self) (srv.py:118); here the
    node subscribes to the OccupancyGrid nav2 publishes on '/local_costmap/costmap' AND to the
    map_msgs/OccupancyGridUpdate patches on '/local_costmap/costmap_updates' (with nav2's default
    always_send_full_costmap: false only patches are sent while the costmap's origin stands still — e.g. while the
    robot waits out a collision stop, srv.py:374-382 — so ignoring them would freeze the obstacle picture)
  * until BOTH a costmap and a footprint have arrived the handler answers with a zero twist instead of solving in free
    space (the reference cannot solve without its Costmap2d either, and objective() crashes without self.footprint)
  * dynamic parameters                                                               srv.py:405-439

``rclpy``, ``neo_srvs2`` and the message packages are imported when this module is imported; they are not part of this
repository's image, so the module is exercised in the tests under the stand-in modules of ``oracle/ros_stubs.py``
(the same ones that drive the unmodified reference for the golden vectors).  There is no CPU path: constructing the node
without a CUDA device raises.

    ros2 run ... python -m neo_mpc_planner2_b200.ros_node --ros-args --params-file navigation.yaml
"""
from __future__ import annotations

import math
import time

import numpy as np
import rclpy
from rclpy.node import Node
from rclpy.parameter import Parameter
from rcl_interfaces.msg import SetParametersResult
from geometry_msgs.msg import PolygonStamped, PoseStamped
from nav_msgs.msg import OccupancyGrid, Path
try:                                      # map_msgs ships with nav2; the stand-in test modules provide it too
    from map_msgs.msg import OccupancyGridUpdate
except ImportError:                       # pragma: no cover
    OccupancyGridUpdate = None
from neo_srvs2.srv import Optimizer

from .abi import ENC_OCCUPANCY, REQUEST_DTYPE
from .server import request_to_msg
from .solver import BatchSolver

# the reference's declarations (srv.py:49-75): name -> code default
REFERENCE_PARAMETERS = dict(
    acc_x_limit=0.5, acc_y_limit=0.5, acc_theta_limit=0.5,
    min_vel_x=-0.5, min_vel_y=-0.5, min_vel_trans=0.5, min_vel_theta=-0.5,
    max_vel_x=0.5, max_vel_y=0.5, max_vel_trans=0.5, max_vel_theta=0.5,
    w_trans=0.5, w_orient=0.5, w_control=0.5, w_terminal=0.5, w_costmap=0.5, w_footprint=2000,
    waiting_time=3.0, low_pass_gain=0.5, opt_tolerance=1e-5, prediction_horizon=0.5, control_steps=3)

# Names cb_params accepts (srv.py:408-436).  In the reference only some of them change the solve afterwards: the
# bounds list is built once from min/max_vel_* (srv.py:125-133) and w_costmap / w_footprint are copied to
# *_scale attributes at start-up (srv.py:96-97) which the callback does not touch.
DYNAMIC_NAMES = ("min_vel_x", "min_vel_y", "min_vel_trans", "min_vel_theta", "max_vel_x", "max_vel_y", "max_vel_trans",
                 "max_vel_theta", "w_trans", "w_orient", "w_control", "w_terminal", "w_costmap", "w_footprint")
EFFECTIVE_IN_REFERENCE = ("max_vel_trans", "w_trans", "w_orient", "w_control", "w_terminal")


def yaw_of(q):
    """euler_from_quaternion yaw (srv.py:176-178)."""
    return math.atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z))


class MpcOptimizationServer(Node):
    def __init__(self, device: int = 0, strict_reference_parameters: bool = True, **solver_knobs):
        super().__init__('mpc_optimization_server')
        for name, default in REFERENCE_PARAMETERS.items():                       # srv.py:49-75
            self.declare_parameter(name, value=default)
        self.params = {name: self.get_parameter(name).value for name in REFERENCE_PARAMETERS}   # srv.py:78-103
        self.strict_reference_parameters = strict_reference_parameters
        self._solver = BatchSolver(self.params, device=device, **solver_knobs)
        self._solver.reserve_instances(1)

        self.srv = self.create_service(Optimizer, 'optimizer', self.optimizer)   # srv.py:105
        self.add_on_set_parameters_callback(self.cb_params)                      # srv.py:106
        self.PubRaysPath = self.create_publisher(Path, 'local_plan', 10)         # srv.py:107
        self.local_plan = Path()
        self.footprint = None                                                    # world-frame polygon, srv.py:154-155
        self.subscription_footprint = self.create_subscription(
            PolygonStamped, '/local_costmap/published_footprint', self.footprint_callback, 10)   # srv.py:140-144
        self.subscription_costmap = self.create_subscription(
            OccupancyGrid, '/local_costmap/costmap', self.costmap_callback, 10)  # stands for Costmap2d(self), srv.py:118
        self.subscription_costmap_updates = None
        if OccupancyGridUpdate is not None:
            self.subscription_costmap_updates = self.create_subscription(
                OccupancyGridUpdate, '/local_costmap/costmap_updates', self.costmap_update_callback, 10)
        self._grid = None                                                        # cached cells of the last full costmap
        self._grid_info = None
        self.costmap_generation = 0
        self.last_time = 0.0                                                     # srv.py:138
        self.last_response = None
        self.solution = None

    # ---- inputs from topics
    def footprint_callback(self, msg):
        self.footprint = msg.polygon                                             # srv.py:154-155

    def costmap_callback(self, msg):
        info = msg.info
        self._grid = np.array(msg.data, dtype=np.int8).reshape(info.height, info.width)
        self._grid_info = (info.resolution, info.origin.position.x, info.origin.position.y)
        self._upload_costmap()

    def costmap_update_callback(self, msg):
        """map_msgs/OccupancyGridUpdate: a rectangular patch (x, y, width, height, data) of the last full grid."""
        if self._grid is None:
            return                                                               # no full grid yet: nothing to patch
        h, w = self._grid.shape
        x0, y0, pw, ph = int(msg.x), int(msg.y), int(msg.width), int(msg.height)
        if x0 < 0 or y0 < 0 or x0 + pw > w or y0 + ph > h or pw * ph != len(msg.data):
            self.get_logger().warn("costmap update outside the cached grid: ignored until the next full costmap")
            return
        self._grid[y0:y0 + ph, x0:x0 + pw] = np.asarray(msg.data, dtype=np.int8).reshape(ph, pw)
        self._upload_costmap()

    def _upload_costmap(self):
        res, ox, oy = self._grid_info
        self._solver.set_costmap(self._grid, res, ox, oy, ENC_OCCUPANCY)
        self.costmap_generation += 1

    def _place_footprint(self, pose):
        """The library takes the polygon in the robot frame and places it at the request's current pose; the reference
        holds the already-placed world-frame polygon.  Expressing the received polygon in the frame of the request's
        pose makes the library test exactly the polygon the reference would (whatever pose it was published at)."""
        if self.footprint is None or not self.footprint.points:
            return
        yaw = yaw_of(pose.orientation)
        c, s = math.cos(yaw), math.sin(yaw)
        xy = []
        for p in self.footprint.points:
            dx, dy = p.x - pose.position.x, p.y - pose.position.y
            xy.append((c * dx + s * dy, -s * dx + c * dy))
        self._solver.set_footprint(np.asarray(xy, dtype=np.float32))

    # ---- the service handler (srv.py:349-403)
    def optimizer(self, request, response):
        current_time = time.time()                                               # srv.py:369-371
        delta_t = current_time - self.last_time
        self.last_time = current_time
        if self._grid is None or self.footprint is None or not self.footprint.points:
            # no obstacle picture yet: standing still is the only safe answer (nothing is solved in free space)
            self.get_logger().warn("optimizer called before a costmap and a footprint were received: zero twist")
            response.output_vel.twist.linear.x = 0.0
            response.output_vel.twist.linear.y = 0.0
            response.output_vel.twist.angular.z = 0.0
            return response
        self._place_footprint(request.current_pose.pose)
        msg = request_to_msg(request, delta_t, instance_id=0)
        out, plan = self._solver.solve_msgs(msg, want_plan=True)
        response.output_vel.twist.linear.x = float(out["vx"][0])                 # srv.py:375-377 / :389-391
        response.output_vel.twist.linear.y = float(out["vy"][0])
        response.output_vel.twist.angular.z = float(out["omega"][0])
        self.last_response = out[0]
        self.solution = plan[0]
        self.publishLocalPlan(plan[0], request.current_pose.pose)                # srv.py:365
        return response

    def publishLocalPlan(self, x, start_pose):
        """srv.py:271-310; the start pose is the request's current_pose instead of a TF lookup (srv.py:274-286)."""
        req = np.zeros(1, REQUEST_DTYPE)
        req["pose_x"], req["pose_y"] = start_pose.position.x, start_pose.position.y
        req["pose_yaw"] = yaw_of(start_pose.orientation)
        poses = self._solver.local_plan(req, np.asarray(x, np.float32)[None, :])[0]
        self.local_plan.poses.clear()                                            # srv.py:272
        stamp = self.get_clock().now().to_msg()
        for p in poses:
            ps = PoseStamped()
            ps.pose.position.x, ps.pose.position.y = float(p["x"]), float(p["y"])
            ps.pose.orientation.z, ps.pose.orientation.w = float(p["qz"]), float(p["qw"])
            ps.header.stamp = stamp
            self.local_plan.poses.append(ps)
        self.local_plan.header.stamp = stamp                                     # srv.py:308-310
        self.local_plan.header.frame_id = "map"
        self.PubRaysPath.publish(self.local_plan)

    # ---- dynamic parameters (srv.py:405-439)
    def cb_params(self, data):
        changes = {}
        for parameter in data:
            if parameter.type_ == Parameter.Type.DOUBLE:
                if parameter.name in DYNAMIC_NAMES:
                    if not self.strict_reference_parameters or parameter.name in EFFECTIVE_IN_REFERENCE:
                        changes[parameter.name] = parameter.value
                else:
                    print("The selected parameter cannot be dynamically changed")
        if changes:
            self.params.update(changes)
            self._solver.set_params(**changes)          # solver knobs (footprint_mode, ...) keep their values
        return SetParametersResult(successful=True)

    def destroy_node(self):
        self._solver.close()
        parent = getattr(super(), "destroy_node", None)
        if parent is not None:
            parent()


def main(args=None):
    rclpy.init(args=args)                                                        # srv.py:441-444
    node = MpcOptimizationServer()
    rclpy.spin(node)


if __name__ == '__main__':
    main()

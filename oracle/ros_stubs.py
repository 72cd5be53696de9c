"""Minimal fake ROS 2 modules so the UNMODIFIED reference file
``/root/reference/neo_mpc_planner2/mpc_optimization_server.py`` can be imported in the build
container (no rclpy / nav2 / neo_srvs2 there).  Used only by ``make_golden.py``; nothing in the
GPU tests, ``smoke()`` or ``bench.py`` needs the reference at run time.

The fakes carry no algorithm: message types are plain dataclasses, the node base class stores
parameters, publishers swallow messages and TF lookups always fail (so ``publishLocalPlan``
returns early at srv.py:279-282).
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types
from dataclasses import dataclass, field


# ----------------------------------------------------------------------------- messages
@dataclass
class Vector3:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0


@dataclass
class Point:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0


@dataclass
class Point32:
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
class Header:
    stamp: object = None
    frame_id: str = ""


@dataclass
class Pose:
    position: Point = field(default_factory=Point)
    orientation: Quaternion = field(default_factory=Quaternion)


@dataclass
class PoseStamped:
    header: Header = field(default_factory=Header)
    pose: Pose = field(default_factory=Pose)


@dataclass
class Twist:
    linear: Vector3 = field(default_factory=Vector3)
    angular: Vector3 = field(default_factory=Vector3)


@dataclass
class TwistStamped:
    header: Header = field(default_factory=Header)
    twist: Twist = field(default_factory=Twist)


@dataclass
class Polygon:
    points: list = field(default_factory=list)


@dataclass
class PolygonStamped:
    header: Header = field(default_factory=Header)
    polygon: Polygon = field(default_factory=Polygon)


@dataclass
class Path:
    header: Header = field(default_factory=Header)
    poses: list = field(default_factory=list)


@dataclass
class OccupancyGrid:
    header: Header = field(default_factory=Header)
    data: list = field(default_factory=list)


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


class Optimizer:
    Request = OptimizerRequest
    Response = OptimizerResponse


@dataclass
class OccupancyGridUpdate:
    header: Header = field(default_factory=Header)
    x: int = 0
    y: int = 0
    width: int = 0
    height: int = 0
    data: list = field(default_factory=list)


@dataclass
class SetParametersResult:
    successful: bool = True
    reason: str = ""


class TransformException(Exception):
    pass


# ----------------------------------------------------------------------------- node
PARAM_OVERRIDES: dict = {}


class _Param:
    def __init__(self, value):
        self.value = value


class _Sink:
    def publish(self, msg):
        pass


class _Logger:
    def info(self, *a, **k):
        pass

    warn = error = debug = info


class _Stamp:
    def to_msg(self):
        return None


class _Clock:
    def now(self):
        return _Stamp()


class Node:
    def __init__(self, name):
        self._name = name
        self._params = {}

    def declare_parameter(self, name, value=None):
        self._params[name] = PARAM_OVERRIDES.get(name, value)

    def get_parameter(self, name):
        return _Param(self._params[name])

    def create_service(self, *a, **k):
        return object()

    def create_publisher(self, *a, **k):
        return _Sink()

    def create_subscription(self, *a, **k):
        return object()

    def add_on_set_parameters_callback(self, cb):
        pass

    def get_logger(self):
        return _Logger()

    def get_clock(self):
        return _Clock()


class _Buffer:
    def lookup_transform(self, *a, **k):
        raise TransformException("no TF in the golden-vector harness")


class _Listener:
    def __init__(self, *a, **k):
        pass


class _PlaceholderCostmap:
    """Replaced by ``oracle.costmap.GridCostmap`` after construction."""

    def __init__(self, node):
        pass


class _ParamType:
    DOUBLE = 3


class Parameter:
    Type = _ParamType


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    rclpy = _mod("rclpy", init=lambda args=None: None, spin=lambda n: None)
    rclpy.time = _mod("rclpy.time", Time=lambda: None)
    rclpy.node = _mod("rclpy.node", Node=Node)
    rclpy.parameter = _mod("rclpy.parameter", Parameter=Parameter)
    _mod("neo_srvs2")
    _mod("neo_srvs2.srv", Optimizer=Optimizer)
    _mod("geometry_msgs")
    _mod("geometry_msgs.msg", TwistStamped=TwistStamped, PoseStamped=PoseStamped, Pose=Pose,
         Polygon=Polygon, PolygonStamped=PolygonStamped, Twist=Twist, Point32=Point32)
    _mod("nav_msgs")
    _mod("nav_msgs.msg", OccupancyGrid=OccupancyGrid, Path=Path)
    _mod("map_msgs")
    _mod("map_msgs.msg", OccupancyGridUpdate=OccupancyGridUpdate)
    _mod("neo_nav2_py_costmap2D")
    _mod("neo_nav2_py_costmap2D.line_iterator", LineIterator=object)
    _mod("neo_nav2_py_costmap2D.costmap", Costmap2d=_PlaceholderCostmap)
    tf2 = _mod("tf2_ros", TransformException=TransformException)
    tf2.buffer = _mod("tf2_ros.buffer", Buffer=_Buffer)
    tf2.transform_listener = _mod("tf2_ros.transform_listener", TransformListener=_Listener)
    _mod("rcl_interfaces")
    _mod("rcl_interfaces.msg", SetParametersResult=SetParametersResult)


def load_reference(path="/root/reference/neo_mpc_planner2/mpc_optimization_server.py"):
    """Load the unmodified reference file by path (its package __init__ imports a missing module)."""
    install()
    spec = importlib.util.spec_from_file_location("ref_mpc_optimization_server", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


REF_PYC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "mpc_optimization_server.pyc.bin")


def load_reference_compiled(path=REF_PYC):
    """Load the byte-compiled, unmodified reference module (oracle/_ref, built by oracle/build_ref.py from the sources
    where they lie under /root/reference): this is what travels to the GPU box, where /root/reference does not exist."""
    install()
    loader = importlib.machinery.SourcelessFileLoader("ref_mpc_optimization_server", path)
    spec = importlib.util.spec_from_loader("ref_mpc_optimization_server", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def make_server(mod, overrides: dict):
    PARAM_OVERRIDES.clear()
    PARAM_OVERRIDES.update(overrides)
    srv = mod.MpcOptimizationServer()
    PARAM_OVERRIDES.clear()
    return srv

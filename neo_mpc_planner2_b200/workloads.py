"""Seeded synthetic workloads for BASELINE.json's configs C1..C5.

These are the inputs of the hot path (request records + a shared costmap).  They are generated in
float64 and rounded ONCE to the float32 request record; both the CUDA path and the oracle then
read the same float32 values, so parity is always measured on identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .abi import REQUEST_DTYPE, STATELESS, README_SAMPLE, ENC_OCCUPANCY

# MPO-700-like rectangular footprint (half sizes 0.4 x 0.3 m).  The reference does not carry the
# robot's footprint (README.md:92 points to an external yaml) — this is a stated assumption.
FOOTPRINT_RECT = [(0.4, 0.3), (-0.4, 0.3), (-0.4, -0.3), (0.4, -0.3)]


@dataclass
class Workload:
    name: str
    params: dict                       # reference parameter names -> values
    requests: np.ndarray               # REQUEST_DTYPE [B]
    cells: np.ndarray | None           # uint8 [H, W] occupancy 0..100, or None (free space)
    resolution: float = 0.05
    origin_x: float = 0.0
    origin_y: float = 0.0
    encoding: int = ENC_OCCUPANCY
    footprint: list = field(default_factory=lambda: list(FOOTPRINT_RECT))

    @property
    def batch(self):
        return len(self.requests)

    @property
    def control_steps(self):
        return int(self.params["control_steps"])

    def algorithmic_bytes_per_solve(self, batch_on_gpu=None):
        """64 B request + 12 B (vx, vy, omega) + costmap read once per launch."""
        b = self.batch if batch_on_gpu is None else batch_on_gpu
        cm = 0 if self.cells is None else self.cells.size
        return 64.0 + 12.0 + cm / float(b)


def quirk_yaw_planar(pose_yaw, goal_yaw):
    """yaw(x=0, y=0, z=sin(pose_yaw/2), w=cos(GOAL_yaw/2)) — the srv.py:213 quirk for planar poses."""
    z = np.sin(np.asarray(pose_yaw, dtype=np.float64) * 0.5)
    w = np.cos(np.asarray(goal_yaw, dtype=np.float64) * 0.5)
    return np.arctan2(2.0 * (w * z), 1.0 - 2.0 * (z * z))


def make_costmap(seed: int, width: int, height: int, resolution: float, n_rect: int,
                 inscribed_radius: float = 0.3, decay: float = 3.0, inflation_radius: float = 1.3):
    """Occupancy grid 0..100: ``n_rect`` axis-aligned lethal rectangles (side U[0.2, 1.0] m) + a
    nav2-style inflation layer: 99 within the inscribed radius, then round(98*exp(-decay*(d-r)))
    out to ``inflation_radius``."""
    from scipy import ndimage

    rng = np.random.default_rng(seed)
    occ = np.zeros((height, width), dtype=bool)
    sx = rng.uniform(0.2, 1.0, n_rect) / resolution
    sy = rng.uniform(0.2, 1.0, n_rect) / resolution
    cx = rng.uniform(0, width, n_rect)
    cy = rng.uniform(0, height, n_rect)
    for k in range(n_rect):
        x0, x1 = int(max(0, cx[k] - sx[k] / 2)), int(min(width, cx[k] + sx[k] / 2 + 1))
        y0, y1 = int(max(0, cy[k] - sy[k] / 2)), int(min(height, cy[k] + sy[k] / 2 + 1))
        occ[y0:y1, x0:x1] = True
    d = ndimage.distance_transform_edt(~occ) * resolution
    cells = np.zeros((height, width), dtype=np.uint8)
    ring = (d > inscribed_radius) & (d <= inflation_radius)
    cells[ring] = np.round(98.0 * np.exp(-decay * (d[ring] - inscribed_radius))).astype(np.uint8)
    cells[d <= inscribed_radius] = 99
    cells[occ] = 100
    return cells


def make_requests(seed: int, batch: int, *, extent_x, extent_y, cells=None, resolution=0.05,
                  origin=(0.0, 0.0), margin=1.0, carrot_range=0.4, control_interval=1.0 / 30.0,
                  stateless=True, carrot_bearings=None, poses=None):
    """Random requests: start pose uniform at least ``margin`` inside the map (and,
    with a costmap, not inside an inscribed/lethal cell), yaw U[-pi, pi]; current velocity
    U[-0.3, 0.3]^3; carrot at ``carrot_range`` in the base frame, bearing U[-pi, pi] (or the given
    bearings), carrot yaw U[-1, 1]; goal anywhere in the map, yaw U[-pi, pi]."""
    rng = np.random.default_rng(seed)
    ox, oy = origin
    lo_x, hi_x = ox + margin, ox + extent_x - margin
    lo_y, hi_y = oy + margin, oy + extent_y - margin
    if poses is None:
        px = np.empty(batch)
        py = np.empty(batch)
        filled = 0
        while filled < batch:
            m = max(1024, int((batch - filled) * 1.6))
            tx = rng.uniform(lo_x, hi_x, m)
            ty = rng.uniform(lo_y, hi_y, m)
            if cells is not None:
                mx = ((tx - ox) / resolution).astype(np.int64)
                my = ((ty - oy) / resolution).astype(np.int64)
                ok = cells[my, mx] < 99
                tx, ty = tx[ok], ty[ok]
            k = min(len(tx), batch - filled)
            px[filled:filled + k] = tx[:k]
            py[filled:filled + k] = ty[:k]
            filled += k
        pyaw = rng.uniform(-math.pi, math.pi, batch)
    else:
        px, py, pyaw = poses
    req = np.zeros(batch, dtype=REQUEST_DTYPE)
    v0 = rng.uniform(-0.3, 0.3, (batch, 3))
    bearing = rng.uniform(-math.pi, math.pi, batch) if carrot_bearings is None else carrot_bearings
    req["vel_x"], req["vel_y"], req["vel_theta"] = v0[:, 0], v0[:, 1], v0[:, 2]
    req["carrot_x"] = carrot_range * np.cos(bearing)
    req["carrot_y"] = carrot_range * np.sin(bearing)
    req["carrot_yaw"] = rng.uniform(-1.0, 1.0, batch)
    req["goal_x"] = rng.uniform(ox + margin, ox + extent_x - margin, batch)
    req["goal_y"] = rng.uniform(oy + margin, oy + extent_y - margin, batch)
    req["goal_yaw"] = rng.uniform(-math.pi, math.pi, batch)
    req["pose_x"], req["pose_y"], req["pose_yaw"] = px, py, pyaw
    # the quirk yaw is computed from the float32-rounded yaws the record actually carries
    req["pose_yaw_objective"] = quirk_yaw_planar(req["pose_yaw"].astype(np.float64),
                                                 req["goal_yaw"].astype(np.float64))
    req["control_interval"] = control_interval
    req["delta_t"] = control_interval
    req["instance_id"] = STATELESS if stateless else np.arange(batch, dtype=np.uint32)
    return req


def _params(**over):
    p = dict(README_SAMPLE)
    p.update(over)
    return p


def kat_request():
    """The single known-answer problem of SURVEY.md §8c (config C1)."""
    req = np.zeros(1, dtype=REQUEST_DTYPE)
    req["carrot_x"], req["carrot_y"], req["carrot_yaw"] = 0.4, 0.1, 0.3
    req["goal_x"], req["goal_y"], req["goal_yaw"] = 3.0, 1.0, 0.5
    req["pose_x"], req["pose_y"], req["pose_yaw"] = 1.0, 2.0, 0.2
    req["pose_yaw_objective"] = quirk_yaw_planar(req["pose_yaw"].astype(np.float64),
                                                 req["goal_yaw"].astype(np.float64))
    req["control_interval"] = 1.0 / 30.0
    req["delta_t"] = 1.0 / 30.0
    req["instance_id"] = STATELESS
    return req


def config(name: str, batch: int | None = None, seed: int | None = None) -> Workload:
    """BASELINE.json configs.  ``batch`` overrides the config's batch size (parity tests use small
    slices of the same distribution; the bench uses the full size)."""
    name = name.lower()
    if name == "c1":
        return Workload("c1", _params(control_steps=3), kat_request(), None)
    if name == "c2":
        w = h = 200
        cells = make_costmap(2, w, h, 0.05, 12)
        req = make_requests(2 if seed is None else seed, batch or 4096, extent_x=10.0, extent_y=10.0,
                            cells=cells, origin=(-5.0, -5.0), margin=1.5)
        return Workload("c2", _params(control_steps=3), req, cells, 0.05, -5.0, -5.0)
    if name in ("c3", "c4"):
        w = h = 1000
        cells = make_costmap(3, w, h, 0.05, 300)
        n = 10 if name == "c3" else 20
        s = (3 if name == "c3" else 4) if seed is None else seed
        b = batch or (65536 if name == "c3" else 1048576)
        req = make_requests(s, b, extent_x=50.0, extent_y=50.0, cells=cells,
                            origin=(-25.0, -25.0), margin=1.5)
        return Workload(name, _params(control_steps=n, w_footprint=2000), req, cells, 0.05, -25.0, -25.0)
    if name == "c5":
        w = h = 2000
        cells = make_costmap(5, w, h, 0.05, 1200)
        n_pose = batch or 100000
        s = 5 if seed is None else seed
        base = make_requests(s, n_pose, extent_x=100.0, extent_y=100.0, cells=cells,
                             origin=(-50.0, -50.0), margin=1.5)
        reqs = np.repeat(base, 8)
        k = np.tile(np.arange(8), n_pose)
        bearing = k * (math.pi / 4.0)
        reqs["carrot_x"] = 0.4 * np.cos(bearing)
        reqs["carrot_y"] = 0.4 * np.sin(bearing)
        return Workload("c5", _params(control_steps=10, w_footprint=2000), reqs, cells, 0.05, -50.0, -50.0)
    raise ValueError(f"unknown config {name!r}")

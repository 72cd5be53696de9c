"""CPU restatement of the plugin's carrot selection (ORACLE / test infrastructure only) — SURVEY.md §8f row N2,
the step immediately BEFORE the MPC solve: reference src/NeoMpcPlanner.cpp ("cpp").

  transformGlobalPlan   cpp:66-135   closest plan pose (min_by), closer_to_goal, window end, pruning, base-frame transform
  getLookAheadDistance  cpp:157-171  lookahead by slow_down_ / closer_to_goal
  getLookAheadPoint     cpp:173-189  first pose at least lookahead_dist away, else the last one
  slow-down hysteresis  cpp:216-236  from |yaw(carrot)| and the footprint cost at the robot pose; 255 -> exception

The C++ plugin cannot be compiled here (no ROS 2), so this restatement is NOT pinned against the reference: "parity
unpinned" for this row.  TF is out of scope: plan and robot pose are in the same frame, and the base-frame transform
is the planar one  x_b = c*dx + s*dy,  y_b = -s*dx + c*dy,  yaw_b = wrap(yaw_plan - yaw_robot).
Footprint cost: nav2 footprintCostAtPose on RAW costmap bytes (0..254, 255 = no information): max byte over all
rasterised footprint edges, a vertex outside the map gives 254.
"""
from __future__ import annotations

import math

import numpy as np

from .costmap import bresenham_cells, ENC_NAV2_RAW, ENC_OCCUPANCY

STATUS_OK = 0
STATUS_EMPTY_WINDOW = 1      # "Resulting plan has 0 poses in it."      (cpp:130-132)
STATUS_COLLISION = 2         # "MPC detected collision!"               (cpp:234-236)


def raw_byte_table(encoding: int) -> np.ndarray:
    """Costmap byte -> nav2 raw cost (0..255).  Occupancy grids are mapped back with the inverse of nav2's
    publisher table: 0->0, 100->254, 99->253, unknown->255, v in 1..98 -> 1 + round((v-1)*251/97)."""
    t = np.zeros(256, dtype=np.int64)
    if encoding == ENC_NAV2_RAW:
        t[:] = np.arange(256)
    elif encoding == ENC_OCCUPANCY:
        t[:] = 255
        t[0] = 0
        for v in range(1, 99):
            t[v] = 1 + int(math.floor((v - 1) * 251.0 / 97.0 + 0.5))
        t[99] = 253
        t[100] = 254
    else:
        raise ValueError(encoding)
    return t


def footprint_raw_cost(costmap, footprint_robot, x, y, yaw):
    """nav2 FootprintCollisionChecker::footprintCostAtPose restated on raw bytes (declared semantics above)."""
    table = raw_byte_table(costmap.encoding)
    c, s = math.cos(yaw), math.sin(yaw)
    cells = []
    for fx, fy in footprint_robot:
        mx, my = costmap.getWorldToMap(x + (fx * c - fy * s), y + (fx * s + fy * c))
        if mx < 0:
            return 254
        cells.append((mx, my))
    worst = 0
    n = len(cells)
    for k in range(n):
        for cx, cy in bresenham_cells(*cells[k], *cells[(k + 1) % n]):
            if 0 <= cx < costmap.width and 0 <= cy < costmap.height:
                v = int(table[costmap.cells[cy, cx]])
            else:
                v = 254
            worst = max(worst, v)
    return worst


def select_carrot(plan, plan_start, robot, slow_down, lookahead_min, lookahead_max, lookahead_close,
                  max_transform_dist, footprint_cost):
    """One tick of the plugin's front half for one robot.  plan: [L,3] (x, y, yaw); robot: (x, y, yaw).
    Returns dict(status, begin, closer_to_goal, carrot_index, carrot=(x_b, y_b, yaw_b), slow_down)."""
    plan = np.asarray(plan, dtype=np.float64)
    rx, ry, ryaw = robot
    L = len(plan)
    d = np.sqrt((plan[plan_start:, 0] - rx) ** 2 + (plan[plan_start:, 1] - ry) ** 2)
    begin = plan_start + int(np.argmin(d))                               # cpp:81-86 (first minimum)
    dg = math.sqrt((plan[L - 1, 0] - rx) ** 2 + (plan[L - 1, 1] - ry) ** 2)
    closer = dg <= lookahead_close                                       # cpp:88-96
    end = L
    for i in range(begin, L):                                            # cpp:98-103
        if math.sqrt((plan[i, 0] - rx) ** 2 + (plan[i, 1] - ry) ** 2) > max_transform_dist:
            end = i
            break
    out = dict(status=STATUS_OK, begin=begin, closer_to_goal=bool(closer), carrot_index=begin,
               carrot=(0.0, 0.0, 0.0), slow_down=bool(slow_down))
    if end == begin:                                                     # cpp:130-132
        out["status"] = STATUS_EMPTY_WINDOW
        return out
    lookahead = lookahead_min                                            # cpp:161-170
    if (not slow_down) or closer:
        lookahead = lookahead_max
        if closer:
            lookahead = lookahead_close
    c, s = math.cos(ryaw), math.sin(ryaw)
    pick = end - 1                                                       # cpp:184-186 (last pose if none far enough)
    for i in range(begin, end):                                          # cpp:178-182
        dx, dy = plan[i, 0] - rx, plan[i, 1] - ry
        xb, yb = c * dx + s * dy, -s * dx + c * dy
        if math.sqrt(xb * xb + yb * yb) >= lookahead:
            pick = i
            break
    dx, dy = plan[pick, 0] - rx, plan[pick, 1] - ry
    dyaw = plan[pick, 2] - ryaw
    yaw_b = math.atan2(math.sin(dyaw), math.cos(dyaw))
    out["carrot_index"] = pick
    out["carrot"] = (c * dx + s * dy, -s * dx + c * dy, yaw_b)
    # slow-down hysteresis (cpp:216-232): with check_pose_up == carrot the inner test of the first branch is never true
    if abs(yaw_b) < 1.0:
        out["slow_down"] = False
    elif abs(yaw_b) >= 1.0 and footprint_cost > 200:
        out["slow_down"] = True
    else:
        out["slow_down"] = False
    if footprint_cost == 255:                                            # cpp:234-236
        out["status"] = STATUS_COLLISION
    return out

"""Declared costmap semantics (ORACLE / test infrastructure only).

The reference reads its costmap through ``neo_nav2_py_costmap2D.costmap.Costmap2d``
(import ``mpc_optimization_server.py:35-36``, ctor ``:118``; calls ``getWorldToMap``
``:246,332``, ``getCost`` ``:247,257,333`` and ``getFootprintCost`` ``:262-263,343``).
That package is third-party, un-vendored and un-versioned (reference ``README.md:22``), so
its behaviour cannot be pinned.  This file DECLARES the semantics both the oracle and the
CUDA path implement; parity at this boundary is "unpinned" (DESIGN.md says so too).

Declared semantics
------------------
* grid: ``uint8`` cells, row-major, ``idx = my * W + mx``; world origin = lower-left corner
  of cell (0,0); square cells of ``resolution`` metres.
* ``getWorldToMap(wx, wy)``: nav2 ``Costmap2D::worldToMap`` — out of bounds when
  ``wx < origin_x`` or ``wy < origin_y``; else ``mx = int((wx-origin_x)/resolution)``
  (truncation), out of bounds when ``mx >= W`` / ``my >= H``.  Out of bounds -> ``(-1, -1)``.
* ``getCost(mx, my)``: normalised cost in [0, 1] through a 256-entry table, because the
  reference compares the result with ``1.0`` (lethal, ``:257,262,343``) and ``0.99``
  (inscribed, ``:338``):
    - ``ENC_OCCUPANCY`` (nav_msgs/OccupancyGrid, what nav2 publishes on
      ``/local_costmap/costmap``: 254->100, 253->99, unknown->-1): ``v/100`` for
      ``0 <= v <= 100``; every other byte (unknown) -> 0.0;
    - ``ENC_NAV2_RAW`` (nav2 ``costmap_raw``): ``v/254`` for ``0 <= v <= 254``; 255
      (NO_INFORMATION) -> 0.0.
  An out-of-bounds cell costs 1.0 (lethal), the nav2 ``FootprintCollisionChecker``
  convention for points that leave the map.
* ``getFootprintCost(polygon)``: nav2 ``FootprintCollisionChecker::footprintCost`` —
  polygon vertices are WORLD coordinates; each vertex -> cell; every polygon edge
  (including last->first) is rasterised with nav2's ``LineIterator`` (Bresenham, both end
  points included); result = max cell cost met; a vertex outside the map -> 1.0.
"""
from __future__ import annotations

import numpy as np

ENC_OCCUPANCY = 0
ENC_NAV2_RAW = 1


def cost_lut(encoding: int) -> np.ndarray:
    """256-entry byte -> normalised cost table (float64)."""
    lut = np.zeros(256, dtype=np.float64)
    if encoding == ENC_OCCUPANCY:
        v = np.arange(0, 101)
        lut[v] = v / 100.0
    elif encoding == ENC_NAV2_RAW:
        v = np.arange(0, 255)
        lut[v] = v / 254.0
    else:
        raise ValueError(f"unknown costmap encoding {encoding}")
    return lut


def bresenham_cells(x0: int, y0: int, x1: int, y1: int):
    """nav2_util::LineIterator restated: yields every cell from (x0,y0) to (x1,y1) inclusive.

    The k-th cell has the closed form used on the GPU:
    ``major = start + k*sign``, ``minor = start + sign*((den//2 + k*numadd)//den)``.
    """
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    sx = 1 if x1 >= x0 else -1
    sy = 1 if y1 >= y0 else -1
    x, y = x0, y0
    if dx >= dy:
        den, num, numadd, npix = dx, dx // 2, dy, dx
        for _ in range(npix + 1):
            yield x, y
            num += numadd
            if num >= den:
                num -= den
                y += sy
            x += sx
    else:
        den, num, numadd, npix = dy, dy // 2, dx, dy
        for _ in range(npix + 1):
            yield x, y
            num += numadd
            if num >= den:
                num -= den
                x += sx
            y += sy


class GridCostmap:
    """Stand-in for ``neo_nav2_py_costmap2D.costmap.Costmap2d`` with the declared semantics."""

    def __init__(self, cells, resolution: float, origin_x: float, origin_y: float,
                 encoding: int = ENC_OCCUPANCY):
        cells = np.ascontiguousarray(cells)
        if cells.dtype == np.int8:
            cells = cells.view(np.uint8)
        if cells.dtype != np.uint8 or cells.ndim != 2:
            raise ValueError("cells must be a 2-D uint8/int8 array [H, W]")
        self.cells = cells
        self.height, self.width = cells.shape
        self.resolution = float(resolution)
        self.origin_x = float(origin_x)
        self.origin_y = float(origin_y)
        self.encoding = int(encoding)
        self.lut = cost_lut(encoding)

    # -- API used by the reference -------------------------------------------------
    def getWorldToMap(self, wx, wy):
        if wx < self.origin_x or wy < self.origin_y:
            return -1, -1
        mx = int((wx - self.origin_x) / self.resolution)
        my = int((wy - self.origin_y) / self.resolution)
        if mx >= self.width or my >= self.height:
            return -1, -1
        return mx, my

    def getCost(self, mx, my):
        if mx < 0 or my < 0 or mx >= self.width or my >= self.height:
            return 1.0
        return float(self.lut[self.cells[my, mx]])

    def getFootprintCost(self, polygon):
        pts = polygon.points if hasattr(polygon, "points") else polygon
        n = len(pts)
        if n == 0:
            return 0.0
        cells = []
        for p in pts:
            px, py = (p.x, p.y) if hasattr(p, "x") else (p[0], p[1])
            mx, my = self.getWorldToMap(px, py)
            if mx < 0:
                return 1.0
            cells.append((mx, my))
        worst = 0.0
        for k in range(n):
            x0, y0 = cells[k]
            x1, y1 = cells[(k + 1) % n]
            for cx, cy in bresenham_cells(x0, y0, x1, y1):
                c = self.getCost(cx, cy)
                if c > worst:
                    worst = c
        return worst

    # -- helpers for vectorised checks --------------------------------------------
    def cost_at_world(self, wx, wy):
        """Vectorised getCost(getWorldToMap(wx, wy)) for numpy arrays."""
        wx = np.asarray(wx, dtype=np.float64)
        wy = np.asarray(wy, dtype=np.float64)
        fx = (wx - self.origin_x) / self.resolution
        fy = (wy - self.origin_y) / self.resolution
        # truncation toward zero == int() in getWorldToMap; negatives are OOB anyway
        mx = np.trunc(fx).astype(np.int64)
        my = np.trunc(fy).astype(np.int64)
        oob = (wx < self.origin_x) | (wy < self.origin_y) | (mx >= self.width) | (my >= self.height)
        mxc = np.clip(mx, 0, self.width - 1)
        myc = np.clip(my, 0, self.height - 1)
        c = self.lut[self.cells[myc, mxc]]
        return np.where(oob, 1.0, c)

    def bilinear_at_world(self, wx, wy):
        """OPT-IN bilinear mode (NEOMPC_COSTMAP_BILINEAR; not the reference's behaviour): normalised cost c and
        lethal indicator l interpolated between the four cell centres around the world points, with their
        derivatives w.r.t. the world coordinates.  Out-of-map cells count as lethal (cost 1.0), as in getCost.
        Returns (c, l, dc_dx, dc_dy, dl_dx, dl_dy), numpy float64 arrays shaped like wx."""
        wx = np.asarray(wx, dtype=np.float64)
        wy = np.asarray(wy, dtype=np.float64)
        gx = (wx - self.origin_x) / self.resolution - 0.5
        gy = (wy - self.origin_y) / self.resolution - 0.5
        i0 = np.floor(gx).astype(np.int64)
        j0 = np.floor(gy).astype(np.int64)
        tx, ty = gx - i0, gy - j0
        lethal_byte = 100 if self.encoding == ENC_OCCUPANCY else 254

        def corner(di, dj):
            cx, cy = i0 + di, j0 + dj
            inb = (cx >= 0) & (cy >= 0) & (cx < self.width) & (cy < self.height)
            b = self.cells[np.clip(cy, 0, self.height - 1), np.clip(cx, 0, self.width - 1)]
            c = np.where(inb, self.lut[b], 1.0)
            l = np.where(inb, (b == lethal_byte).astype(np.float64), 1.0)
            return c, l
        (c00, l00), (c10, l10), (c01, l01), (c11, l11) = corner(0, 0), corner(1, 0), corner(0, 1), corner(1, 1)

        def lerp2(v00, v10, v01, v11):
            v0 = v00 + tx * (v10 - v00)
            v1 = v01 + tx * (v11 - v01)
            v = v0 + ty * (v1 - v0)
            dvx = ((v10 - v00) + ty * ((v11 - v01) - (v10 - v00))) / self.resolution
            dvy = (v1 - v0) / self.resolution
            return v, dvx, dvy
        c, dcx, dcy = lerp2(c00, c10, c01, c11)
        l, dlx, dly = lerp2(l00, l10, l01, l11)
        return c, l, dcx, dcy, dlx, dly

    def edge_distance_cells(self, wx, wy):
        """Distance (in cells) of world points to the nearest cell edge — used by parity
        tests to exclude samples where fp32 vs fp64 rounding may flip the cell index."""
        fx = (np.asarray(wx, dtype=np.float64) - self.origin_x) / self.resolution
        fy = (np.asarray(wy, dtype=np.float64) - self.origin_y) / self.resolution
        dx = np.abs(fx - np.round(fx))
        dy = np.abs(fy - np.round(fy))
        return np.minimum(dx, dy)


class FreeSpaceCostmap:
    """Costmap fake that is free everywhere (BASELINE.json config C1: 'no costmap term')."""
    encoding = ENC_OCCUPANCY

    def getWorldToMap(self, wx, wy):
        return 0, 0

    def getCost(self, mx, my):
        return 0.0

    def getFootprintCost(self, polygon):
        return 0.0

"""Drop-in for the reference map object (src/planner/scripts/map_server/esdf.py, class ESDF): same
constructor, same `occupancy_map_cb(msg)` subscriber callback, same attributes (esdf_map, esdf_grad_x,
esdf_grad_y, occupancy_2d, map_resolution, map_width, map_height, map_origin) and the same query methods --
but the distance field is built on the B200 (exact EDT kernels, bit-identical to scipy/numpy; see
csrc/map_kernels.cuh) and stays resident there for the optimizer. No CPU fallback."""
from __future__ import annotations

import itertools
import threading

import numpy as np

from . import lib

SAFE_DIS = 0.5      # module constant of the reference (ESDF:4), used by has_collision only

_version = itertools.count(1)


class ESDF:
    def __init__(self, device: int = 0):
        self._device = device
        self._handle = None          # private handle used for the build; planners upload into their own slot
        self._lock = threading.Lock()
        self.version = 0             # bumped on every map message: planners re-upload when it changes

    def _h(self):
        if self._handle is None:
            from .worlds import YamlConfig
            self._handle = lib.Handle(YamlConfig(), self._device, 1)
        return self._handle

    def occupancy_map_cb(self, map):
        """ESDF:11-33. `map` is a nav_msgs/OccupancyGrid (or any object with .data and
        .info.{resolution,width,height,origin.position})."""
        raw = np.asarray(map.data)
        H, W = int(map.info.height), int(map.info.width)
        occ = np.where(raw == 100, 100, 0).astype(np.int8).reshape(H, W)       # ESDF:23: unknown is free
        res = map.info.resolution
        origin = map.info.origin.position
        h = self._h()
        with self._lock:                                                        # snapshot per message
            h.set_map_occupancy(0, H, W, float(res), float(origin.x), float(origin.y), occ)
            esdf, gx, gy = h.get_map(0, H, W)
            self.map_resolution = res
            self.map_width = W
            self.map_height = H
            self.map_origin = origin
            self.occupancy_2d = (occ == 100).astype(np.int64)
            self.esdf_map, self.esdf_grad_x, self.esdf_grad_y = esdf, gx, gy
            self.version = next(_version)

    # ---- point queries: host-side reads of the arrays the device built (ESDF:35-82) -----------------------
    def _cell(self, pos):
        r = int((pos[1] - self.map_origin.y) / self.map_resolution)
        c = int((pos[0] - self.map_origin.x) / self.map_resolution)
        if r < 0 or r >= self.map_height or c < 0 or c >= self.map_width:
            return None
        return r, c

    def is_occuiped(self, pos):
        rc = self._cell(pos)
        return False if rc is None else self.occupancy_2d[rc]

    def has_collision(self, pos):
        return self.get_edt_dis(pos) < SAFE_DIS

    def get_edt_dis(self, pos):
        rc = self._cell(pos)
        return 10000 if rc is None else self.esdf_map[rc]

    def get_edt_grad(self, pos):
        rc = self._cell(pos)
        return [0, 0] if rc is None else [self.esdf_grad_x[rc], self.esdf_grad_y[rc]]

    def query_batch(self, xy):
        """Batched get_edt_dis/get_edt_grad on the device: (idx (n,2), dis (n), grad (n,2))."""
        return self._h().query_map(0, xy)

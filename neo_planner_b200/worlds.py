"""Seeded synthetic random-pillar worlds and start/goal problems.

The reference has no benchmark inputs; its world distribution is the Gazebo generator in
``src/simulator/scripts/generate_worlds.py:100-146`` with ``generator_config.yaml:1-16``
(10/15/20 box pillars, side U(0.5,1.5) m, centre x~U(3,27), y~U(-5,5), 1.8 m clearance).
This module restates that distribution directly as a 2-D occupancy grid (what the planner
sees after octomap_server's projection, ``map_server_onboard.launch:9-34``) so that tests,
``bench.py`` and the golden generator consume bit-identical inputs everywhere.

Pure numpy + scipy.ndimage, CPU only. Nothing here is on the product compute path: the
distance field used to *place* start/goal points is computed with scipy on the host so the
problem list does not depend on the code under test.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace

import numpy as np
from scipy import ndimage


@dataclass
class World:
    world_id: int
    H: int
    W: int
    res: float
    ox: float
    oy: float
    occ: np.ndarray  # int8 (H, W); 100 = occupied, 0 = free (nav_msgs/OccupancyGrid values)
    pillars: np.ndarray  # (P, 4): cx, cy, sx, sy

    def occupancy_msg(self):
        """Duck-typed nav_msgs/OccupancyGrid with exactly the fields ESDF:16-20 reads."""
        info = SimpleNamespace(
            resolution=self.res, width=self.W, height=self.H,
            origin=SimpleNamespace(position=SimpleNamespace(x=self.ox, y=self.oy, z=0.0)))
        return SimpleNamespace(data=self.occ.reshape(-1).tolist(), info=info)

    def host_edt(self) -> np.ndarray:
        return ndimage.distance_transform_edt(1 - (self.occ == 100)) * self.res


def make_world(world_id: int, dense: bool = False) -> World:
    """Config A world: 300x300 cells @0.1 m, origin (0,-15). ``dense``: config-5 world,
    1200x1200 @0.05 m (60x60 m), 4x the pillars, y spread widened to +-12 m."""
    rng = np.random.default_rng(world_id)
    if dense:
        H = W = 1200
        res, ox, oy = 0.05, 0.0, -30.0
        n_pillars = 4 * int(rng.choice([10, 15, 20]))
        xr, yr = (3.0, 57.0), (-12.0, 12.0)
    else:
        H = W = 300
        res, ox, oy = 0.1, 0.0, -15.0
        n_pillars = int(rng.choice([10, 15, 20]))
        xr, yr = (3.0, 27.0), (-5.0, 5.0)
    clearance = 1.8
    pillars = []
    for _ in range(n_pillars):
        sx, sy = rng.uniform(0.5, 1.5, size=2)
        for _try in range(10000):
            cx = rng.uniform(*xr)
            cy = rng.uniform(*yr)
            ok = True
            for (px, py, psx, psy) in pillars:
                if abs(cx - px) < (sx + psx) / 2 + clearance and abs(cy - py) < (sy + psy) / 2 + clearance:
                    ok = False
                    break
            if ok:
                break
        else:  # could not place with clearance: drop this pillar
            continue
        pillars.append((cx, cy, sx, sy))
    occ = np.zeros((H, W), dtype=np.int8)
    for (cx, cy, sx, sy) in pillars:
        c0 = int((cx - sx / 2 - ox) / res)
        c1 = int((cx + sx / 2 - ox) / res)
        r0 = int((cy - sy / 2 - oy) / res)
        r1 = int((cy + sy / 2 - oy) / res)
        occ[max(r0, 0):min(r1, H - 1) + 1, max(c0, 0):min(c1, W - 1) + 1] = 100
    return World(world_id, H, W, res, ox, oy, occ, np.array(pillars, dtype=np.float64))


def make_problems(world: World, count: int, M: int = 3, v_max: float = 1.0, safe_dis: float = 0.7,
                  first: int = 0, edt: np.ndarray | None = None):
    """Problems ``first .. first+count-1`` of ``world``.

    Returns ``head (count, 2, 2)`` and ``tail (count, 2, 2)`` [pos; vel] as the ROS node builds
    them (NODE:87, NODE:480-481: two rows, acceleration row absent -> zero-padded, EP:176-182).
    Start points are drawn in the pillar band so that the straight line meets obstacles often.
    """
    if edt is None:
        edt = world.host_edt()
    L = 5.0 * M / 3.0
    margin = safe_dis + 0.1
    head = np.zeros((count, 2, 2))
    tail = np.zeros((count, 2, 2))
    x_hi = world.ox + world.W * world.res - L - 1.0
    y_half = 6.5 if world.H == 300 else 13.5

    def dist(p):
        r = int((p[1] - world.oy) / world.res)
        c = int((p[0] - world.ox) / world.res)
        if r < 0 or r >= world.H or c < 0 or c >= world.W:
            return -1.0
        return edt[r, c]

    for i in range(count):
        k = first + i
        rng = np.random.default_rng(1_000_003 * world.world_id + k)
        while True:
            p0 = np.array([rng.uniform(world.ox + 0.5, x_hi), rng.uniform(-y_half, y_half)])
            if dist(p0) <= margin:
                continue
            th = rng.uniform(-0.6, 0.6)
            d = np.array([np.cos(th), np.sin(th)])
            p1 = p0 + L * d
            if dist(p1) <= margin:
                continue
            break
        head[i, 0] = p0
        head[i, 1] = 0.5 * d
        tail[i, 0] = p1
        tail[i, 1] = 0.8 * v_max * d
    return head, tail


class YamlConfig:
    """Parameter set shipped in src/planner/launch/config/planner_config.yaml:2-13 (the values the
    deployed system runs with; NODE:32-46 loads them into PlannerConfig)."""

    def __init__(self):
        self.v_max = 1.0
        self.T_min = 0.5
        self.T_max = 5.0
        self.safe_dis = 0.7
        self.delta_t = 0.1
        self.weights = [1, 1, 1, 10000]
        self.init_wpts_mode = 'fixed'
        self.init_seg_len = 2.0
        self.init_wpts_num = 2
        self.init_T = 2.5
        self.collision_cost_tol = 5
        self.opt_tol = 1e-2


class LibraryDefaultConfig:
    """Library defaults of the reference optimizer (EP:12-25) - second parity parameter set."""

    def __init__(self):
        self.v_max = 10.0
        self.T_min = 2.0
        self.T_max = 20.0
        self.safe_dis = 0.5
        self.delta_t = 0.1
        self.weights = [1.0, 1.0, 0.001, 10000]
        self.init_wpts_mode = 'fixed'
        self.init_seg_len = 2.0
        self.init_wpts_num = 2
        self.init_T = 2.0
        self.collision_cost_tol = 10
        self.opt_tol = 1e-4

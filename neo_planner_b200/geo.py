"""Geometric initializer: host-side mirror of the reference's
src/planner/scripts/traj_planner/astar_planner.py (AP) and geo_planner.py (GEO) on top of libneoopt.so.

* ``AstarPlanner``    -- drop-in: ``plan(map, start_pos, target_pos) -> [[x, y], ...]`` (AP:22-103).
* ``GeoPlanner``      -- drop-in: ``geo_traj_plan(map, plan_init_state, target_state)``, ``prune_path_nodes``,
  ``seg_feasible_check`` (GEO:19-101); inherits the device-backed MinJerkPlanner.
* ``BatchGeoPlanner`` -- the same for B start/target pairs per call: one ``k_astar`` launch (search + pruning, one warp per
  pair) feeding one ``k_optimize`` launch.

The search, the pruning and the optimisation all run on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import lib
from .planner import BatchPlanner, MinJerkPlanner, _MapCache

_STATUS_TEXT = {lib.ASTAR_START_OUTSIDE: 'start position is outside the search grid',
                lib.ASTAR_LIMIT: 'search limit reached'}


def _check(status):
    for st in np.unique(status):
        if int(st) in _STATUS_TEXT:
            raise ValueError(_STATUS_TEXT[int(st)])


def geo_times(cfg, B=None):
    """GEO:30-32: init_T for every piece, first and last stretched by 1.5 (two interior waypoints -> three pieces)."""
    ts = cfg.init_T * np.ones(3)
    ts[0] *= 1.5
    ts[-1] *= 1.5
    return ts if B is None else np.tile(ts, (B, 1))


class AstarPlanner:
    """AP:6-160. The open-set scan, neighbour tests and path extraction run in k_astar (csrc/astar_warp.cuh)."""

    def __init__(self, device: int = 0, handle=None, maps=None):
        self._device = device
        self._h = handle                 # a GeoPlanner shares its handle and its map cache with its A* planner
        self._maps = maps if maps is not None else (None if handle is None else _MapCache(handle))
        self.last = None

    def _handle(self):
        if self._h is None:
            from .planner import DefaultConfig
            self._h = lib.Handle(DefaultConfig(), self._device, 1)
            self._maps = _MapCache(self._h)
        return self._h

    def search(self, map, start_pos, target_pos):
        h = self._handle()
        self._maps.ensure(0, map)
        start = np.asarray(start_pos, dtype=np.float64)[:2].reshape(1, 2)
        target = np.asarray(target_pos, dtype=np.float64)[:2].reshape(1, 2)
        out = h.astar(start, target)                       # lengths first, then the nodes
        _check(out['status'])
        out = h.astar(start, target, max_path=max(int(out['path_len'][0]), 1))
        self.last = out
        return out

    def plan(self, map, start_pos, target_pos):
        out = self.search(map, start_pos, target_pos)
        if out['status'][0] == lib.ASTAR_EXHAUSTED:
            print("Open set is empty, no path found")      # AP:59
        return [list(p) for p in out['path'][0, :out['path_len'][0]]]


class GeoPlanner(MinJerkPlanner):
    """GEO:13-101."""

    def __init__(self, planner_config=None, device: int = 0):
        super().__init__(planner_config, device)
        self.astar_planner = AstarPlanner(device, self._handle(), self._maps)
        self.int_wpts_num = 2

    def geo_traj_plan(self, map, plan_init_state, target_state):
        start_pos = np.asarray(plan_init_state.global_pos)[:2]
        target_state = np.asarray(target_state, dtype=np.float64)
        out = self.astar_planner.search(map, start_pos, target_state[0])
        self.path_pruned = [list(p) for p in out['pruned'][0]]
        int_wpts = out['pruned'][0, 1:3].T.copy()          # GEO:29
        ts = geo_times(self)
        drone_state_2d = np.array([np.asarray(plan_init_state.global_pos)[:2], np.asarray(plan_init_state.global_vel)[:2]])
        self.warm_start_plan(map, drone_state_2d, target_state, int_wpts, ts)

    def seg_feasible_check(self, map, head_pos, tail_pos):
        """GEO:37-55: <= 0.1 m samples along the segment, 0.4 m clearance at every one (distances read on the device with
        neo_query_map)."""
        x0, y0, x1, y1 = head_pos[0], head_pos[1], tail_pos[0], tail_pos[1]
        n = int(np.ceil(max(abs(x1 - x0), abs(y1 - y0)) / 0.1)) + 1
        pts = np.stack((np.linspace(x0, x1, n), np.linspace(y0, y1, n)), axis=1)
        self._handle(); self._maps.ensure(0, map)
        _, dis, _ = self._h.query_map(0, pts)
        return not bool((dis < 0.4).any())

    def prune_path_nodes(self, map, path):
        """GEO:57-101 for an arbitrary node list (geo_traj_plan itself prunes inside k_astar): greedy line-of-sight key
        nodes, then exactly four of them."""
        nodes = [list(p) for p in path]
        keys, head, tail = [0], 0, 1
        while tail < len(nodes):
            while tail < len(nodes) and (self.seg_feasible_check(map, nodes[head], nodes[tail]) or tail - head == 1):
                tail += 1
            keys.append(tail - 1)
            head = tail - 1
        return [nodes[i] for i in four_key_indices(keys)]


def four_key_indices(key_index):
    """GEO:78-95."""
    n = len(key_index)
    if n == 2:
        return [int(v) for v in np.linspace(key_index[0], key_index[-1], 4).astype(int)]
    if n == 3:
        a, b, c = key_index
        return [a, int((a + b) / 2), b, c] if b - a > c - b else [a, b, int((b + c) / 2), c]
    if n == 4:
        return list(key_index)
    left, right = 1 / 3 * key_index[-1], 2 / 3 * key_index[-1]
    return [key_index[0], min(key_index, key=lambda v: abs(v - left)), min(key_index, key=lambda v: abs(v - right)),
            key_index[-1]]


class BatchGeoPlanner(BatchPlanner):
    """geo_traj_plan for B problems: k_astar (search + pruning) then warm_start_plan from the two middle key nodes."""

    def geo_guess(self, head, tail, map_ids=None, max_path=0, max_closed=0):
        head = np.asarray(head, dtype=np.float64); tail = np.asarray(tail, dtype=np.float64)
        out = self.handle.astar(head[:, 0], tail[:, 0], map_ids, max_path, max_closed)
        _check(out['status'])
        out['int_wpts'] = np.ascontiguousarray(out['pruned'][:, 1:3].transpose(0, 2, 1))     # (B,2,2): rows x, y
        out['ts'] = geo_times(self.cfg, head.shape[0])
        return out

    def geo_plan(self, head, tail, map_ids=None, rng=None, max_attempts=5):
        g = self.geo_guess(head, tail, map_ids)
        res = self.warm_start_plan(head, tail, g['int_wpts'], g['ts'], map_ids, rng, max_attempts)
        res['astar_status'] = g['status']; res['path_len'] = g['path_len']; res['closed'] = g['closed']
        res['pruned'] = g['pruned']
        return res

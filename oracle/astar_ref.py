"""astar_ref.py -- TEST INFRASTRUCTURE ONLY (not shipped, not on the product path).

Python restatement of the reference's geometric initializer:
  AP  = src/planner/scripts/traj_planner/astar_planner.py   (grid A*, 8-connected, Euclidean heuristic)
  GEO = src/planner/scripts/traj_planner/geo_planner.py     (path pruning to two interior waypoints + warm start)
relative to /root/reference. Parity pin: tests/golden/geo_M3.npz, written by oracle/gen_golden.py from the unmodified
reference (matplotlib, which AP imports only for an unused plot helper, is stubbed there); gen_golden.py asserts that
this module reproduces every path, key index and pruned waypoint exactly.

What has to be restated exactly (results are compared bit for bit):
  * the search grid is the map grown by int(10 m / res) cells in width and height with the origin moved by -5 m (AP:37-41);
    node positions are origin + index * res (AP:117), i.e. cell corners, and are tested with map.has_collision (AP:132);
  * the open set is a dict scanned with min(): the first node in insertion order wins ties on g + hypot (AP:62); a node
    keeps its place in that order when a cheaper parent replaces it (AP:94-95); expansion order = AP:107-114;
  * the path is [target cell] + closed parents, reversed (AP:143-151); an exhausted search returns [target cell] only;
  * pruning walks the path with a 0.1 m line check against 0.4 m clearance (GEO:41-55) and then picks exactly four key
    nodes (GEO:78-95); the two middle ones become int_wpts (GEO:29).
"""
from __future__ import annotations

import heapq
import math

import numpy as np

EXPAND = 10.0                                   # AP:37
MOVES = ((1, 0, 1), (0, 1, 1), (-1, 0, 1), (0, -1, 1),
         (-1, -1, math.sqrt(2)), (-1, 1, math.sqrt(2)), (1, -1, math.sqrt(2)), (1, 1, math.sqrt(2)))   # AP:107-114


class SearchGrid:
    """AP:31-41: the map's grid, enlarged so that a target outside the map can still be reached."""

    def __init__(self, gmap):
        self.map = gmap
        self.res = gmap.res
        self.W = gmap.W + int(EXPAND / self.res)
        self.H = gmap.H + int(EXPAND / self.res)
        self.ox = gmap.ox - EXPAND / 2
        self.oy = gmap.oy - EXPAND / 2

    def cell_of(self, px, py):                                   # AP:119-120
        return int((px - self.ox) / self.res), int((py - self.oy) / self.res)

    def pos_of(self, ix, iy):                                    # AP:116-117
        return self.ox + ix * self.res, self.oy + iy * self.res

    def usable(self, ix, iy):                                    # AP:134-141
        if ix < 0 or ix >= self.W or iy < 0 or iy >= self.H:
            return False
        return not self.map.has_collision(self.pos_of(ix, iy))


def astar(gmap, start_pos, target_pos, max_pops=None):
    """AP:22-103. Returns (path [[x, y], ...], found, n_closed). The open set's `first minimum in insertion order`
    rule is kept with a heap keyed (f, insertion number): a node is inserted into the dict at most once (closed nodes
    never reopen, AP:83), so its insertion number is a constant and the dict scan equals the lexicographic minimum."""
    g = SearchGrid(gmap)
    sx, sy = g.cell_of(start_pos[0], start_pos[1])
    tx, ty = g.cell_of(target_pos[0], target_pos[1])
    key = lambda ix, iy: ix + iy * g.W                           # noqa: E731  (AP:124)
    cost = {key(sx, sy): 0.0}
    parent = {key(sx, sy): -1}
    xy = {key(sx, sy): (sx, sy)}
    number = {key(sx, sy): 0}
    closed = {}
    heap = [(0.0 + math.hypot(sx - tx, sy - ty), 0, key(sx, sy))]
    inserted = 1
    t_parent, found = -1, False
    while True:
        cur = None
        while heap:
            f, num, k = heapq.heappop(heap)
            if k in cost and k not in closed and f == cost[k] + math.hypot(xy[k][0] - tx, xy[k][1] - ty):
                cur = k
                break
        if cur is None:                                          # AP:58-60
            break
        cx, cy = xy[cur]
        if cx == tx and cy == ty:                                # AP:66-69
            t_parent, found = parent[cur], True
            break
        closed[cur] = True                                       # AP:72-73
        if max_pops is not None and len(closed) > max_pops:
            raise RuntimeError('search limit')
        for mx, my, mc in MOVES:                                 # AP:76-95
            nx, ny = cx + mx, cy + my
            k = key(nx, ny)
            if k in closed:
                continue
            if not g.usable(nx, ny):
                continue
            c = cost[cur] + mc
            if k not in cost:
                cost[k] = c; parent[k] = cur; xy[k] = (nx, ny); number[k] = inserted
                inserted += 1
                heapq.heappush(heap, (c + math.hypot(nx - tx, ny - ty), number[k], k))
            elif cost[k] > c:
                cost[k] = c; parent[k] = cur
                heapq.heappush(heap, (c + math.hypot(nx - tx, ny - ty), number[k], k))
    path = [list(g.pos_of(tx, ty))]                              # AP:143-151
    p = t_parent
    while p != -1:
        path.append(list(g.pos_of(*xy[p])))
        p = parent[p]
    return path[::-1], found, len(closed)


def astar_plain(gmap, start_pos, target_pos):
    """The same search written the slow way the reference does it (dict + min over all open nodes); used by the tests on
    small cases to check that the heap formulation above selects the same nodes."""
    g = SearchGrid(gmap)
    sx, sy = g.cell_of(start_pos[0], start_pos[1])
    tx, ty = g.cell_of(target_pos[0], target_pos[1])
    open_ = {sx + sy * g.W: (sx, sy, 0.0, -1)}
    closed = {}
    t_parent = -1
    while open_:
        cur = min(open_, key=lambda o: open_[o][2] + math.hypot(open_[o][0] - tx, open_[o][1] - ty))
        cx, cy, cc, cp = open_[cur]
        if cx == tx and cy == ty:
            t_parent = cp
            break
        del open_[cur]
        closed[cur] = (cx, cy, cc, cp)
        for mx, my, mc in MOVES:
            nx, ny = cx + mx, cy + my
            k = nx + ny * g.W
            if k in closed or not g.usable(nx, ny):
                continue
            if k not in open_ or open_[k][2] > cc + mc:
                open_[k] = (nx, ny, cc + mc, cur)
    path = [list(g.pos_of(tx, ty))]
    p = t_parent
    while p != -1:
        path.append(list(g.pos_of(closed[p][0], closed[p][1])))
        p = closed[p][3]
    return path[::-1]


def segment_clear(gmap, a, b):
    """GEO:41-55: sample the segment every <= 0.1 m (Chebyshev) and require 0.4 m clearance everywhere."""
    n = math.ceil(max(abs(b[0] - a[0]), abs(b[1] - a[1])) / 0.1) + 1
    xs = np.linspace(a[0], b[0], n)
    ys = np.linspace(a[1], b[1], n)
    for i in range(n):
        if gmap.get_edt_dis([xs[i], ys[i]]) < 0.4:
            return False
    return True


def key_nodes(gmap, path):
    """GEO:61-75: greedy line-of-sight shortcutting; returns the indices of the kept path nodes."""
    keys = [0]
    head, tail = 0, 1
    while tail < len(path):
        while segment_clear(gmap, path[head], path[tail]) or tail - head == 1:
            tail += 1
            if tail == len(path):
                break
        keys.append(tail - 1)
        head = tail - 1
    return keys


def four_of(keys):
    """GEO:78-95: reduce / pad the key indices to exactly four."""
    n = len(keys)
    if n == 2:
        return [int(v) for v in np.linspace(keys[0], keys[-1], 4).astype(int)]
    if n == 3:
        if keys[1] - keys[0] > keys[2] - keys[1]:
            return [keys[0], int((keys[0] + keys[1]) / 2), keys[1], keys[2]]
        return [keys[0], keys[1], int((keys[1] + keys[2]) / 2), keys[2]]
    if n == 4:
        return list(keys)
    left = 1 / 3 * keys[-1]
    right = 2 / 3 * keys[-1]
    return [keys[0], min(keys, key=lambda v: abs(v - left)), min(keys, key=lambda v: abs(v - right)), keys[-1]]


def prune(gmap, path):
    """GEO:57-101 -> (four path nodes, their indices, all key indices)."""
    keys = key_nodes(gmap, path)
    pick = four_of(keys)
    return [path[i] for i in pick], pick, keys


def geo_guess(gmap, start_pos, target_pos, init_T):
    """GEO:19-33: A* path -> pruned nodes -> int_wpts (2,2), ts (3,)."""
    path, found, _ = astar(gmap, start_pos, target_pos)
    four, _, _ = prune(gmap, path)
    int_wpts = np.array(four[1:3]).T
    ts = init_T * np.ones((3,))
    ts[0] *= 1.5
    ts[-1] *= 1.5
    return int_wpts, ts, path

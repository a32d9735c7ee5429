"""GPU tests of the geometric initializer: k_astar through the C ABI (neo_astar) and the drop-in
classes geo.AstarPlanner / geo.GeoPlanner / geo.BatchGeoPlanner, against golden results of the unmodified reference
(tests/golden/geo_M3.npz) and against oracle/astar_ref.py on unseen problems. Paths, key nodes and expansion counts are
integer/grid work: compared exactly."""
from types import SimpleNamespace as NS

import numpy as np
import pytest

from neo_planner_b200 import lib
from neo_planner_b200.esdf import ESDF
from neo_planner_b200.geo import AstarPlanner, BatchGeoPlanner, GeoPlanner
from neo_planner_b200.worlds import make_problems, make_world, YamlConfig
from oracle import astar_ref, minco_ref

pytestmark = pytest.mark.gpu


def _grid(w):
    return minco_ref.GridMap(w.occ, w.H, w.W, w.res, w.ox, w.oy)


def test_astar_matches_reference_golden(golden):
    g = golden('geo_M3.npz')
    off = np.concatenate(([0], np.cumsum(g['path_len'])))
    bp = BatchGeoPlanner(YamlConfig(), max_maps=1)
    for wid, dn in sorted(set(zip(g['world_id'].tolist(), g['dense'].tolist()))):
        sel = np.nonzero((g['world_id'] == wid) & (g['dense'] == dn))[0]
        bp.set_map(make_world(wid, dense=bool(dn)))
        out = bp.handle.astar(g['head'][sel, 0], g['tail'][sel, 0], max_path=int(g['path_len'][sel].max()))
        assert np.all(out['status'] == lib.ASTAR_FOUND)
        assert np.array_equal(out['path_len'], g['path_len'][sel])
        assert np.array_equal(out['pruned'], g['pruned'][sel])
        for j, i in enumerate(sel):
            assert np.array_equal(out['path'][j, :g['path_len'][i]], g['path'][off[i]:off[i + 1]]), (wid, i)


def test_astar_matches_oracle_on_unseen_problems():
    """Fresh worlds and problems (short and long targets): path, four key nodes and the number of expanded nodes."""
    bp = BatchGeoPlanner(YamlConfig(), max_maps=2)
    total = 0
    for slot, (wid, Mlen, n) in enumerate(((7, 3, 96), (9, 8, 48))):
        w = make_world(wid)
        bp.set_map(w, slot)
        gm = _grid(w)
        head, tail = make_problems(w, n, M=Mlen)
        ids = np.full(n, slot, np.int32)
        out = bp.handle.astar(head[:, 0], tail[:, 0], ids, max_path=1024)
        again = bp.handle.astar(head[::-1, 0], tail[::-1, 0], ids, max_path=1024)      # scratch is clean, order-independent
        assert np.array_equal(again['pruned'][::-1], out['pruned']) and np.array_equal(again['closed'][::-1], out['closed'])
        for i in range(n):
            path, found, nclosed = astar_ref.astar(gm, head[i, 0], tail[i, 0])
            four, _, _ = astar_ref.prune(gm, path)
            assert found and out['status'][i] == lib.ASTAR_FOUND and out['path_len'][i] == len(path)
            assert np.array_equal(out['path'][i, :len(path)], np.array(path)), (wid, i)
            assert np.array_equal(out['pruned'][i], np.array(four)), (wid, i)
            assert out['closed'][i] == nclosed
            total += 1
    assert total == 144


def test_astar_edge_cases(golden):
    g = golden('geo_M3.npz')
    occ = g['tiny_occ']
    tiny = NS(occ=occ, H=12, W=16, res=1.0, ox=0.0, oy=0.0)
    bp = BatchGeoPlanner(YamlConfig(), max_maps=1)
    bp.set_map(tiny)
    h = bp.handle
    # unreachable target: the reference exhausts the grid and returns the target cell alone (AP:58-60)
    out = h.astar([[2.5, 2.5]], [[10.5, 6.5]], max_path=8)
    assert out['status'][0] == lib.ASTAR_EXHAUSTED and out['path_len'][0] == 1 and out['closed'][0] == int(g['lost_closed'])
    assert np.array_equal(out['path'][0, :1], g['lost_path']) and np.array_equal(out['pruned'][0], g['lost_pruned'])
    # start and target in the same cell; start off the enlarged grid; expansion limit; target off the grid
    out = h.astar([[2.5, 2.5], [-20.0, 0.0], [2.5, 2.5], [2.5, 2.5]], [[2.6, 2.7], [3.0, 3.0], [10.5, 6.5], [2.5, 80.0]],
                  max_closed=10)
    assert out['status'].tolist() == [lib.ASTAR_FOUND, lib.ASTAR_START_OUTSIDE, lib.ASTAR_LIMIT, lib.ASTAR_LIMIT]
    assert out['path_len'].tolist() == [1, 0, 0, 0] and np.array_equal(out['pruned'][0], np.tile([[2.0, 2.0]], (4, 1)))
    gm = minco_ref.GridMap(occ, 12, 16, 1.0, 0.0, 0.0)
    path, found, nclosed = astar_ref.astar(gm, [2.5, 2.5], [2.5, 80.0])
    out = h.astar([[2.5, 2.5]], [[2.5, 80.0]], max_path=4)
    assert not found and out['status'][0] == lib.ASTAR_EXHAUSTED and out['closed'][0] == nclosed
    assert np.array_equal(out['path'][0, :1], np.array(path))
    # empty batch, truncated path buffer
    assert h.astar(np.zeros((0, 2)), np.zeros((0, 2)))['pruned'].shape == (0, 4, 2)
    full = h.astar([[2.5, 2.5]], [[13.5, 9.5]], max_path=64)
    cut = h.astar([[2.5, 2.5]], [[13.5, 9.5]], max_path=3)
    assert cut['path_len'][0] == full['path_len'][0] > 3 and np.array_equal(cut['path'][0], full['path'][0, :3])
    with pytest.raises(ValueError):
        bp.geo_guess(np.array([[[-20.0, 0.0], [0, 0]]]), np.array([[[3.0, 3.0], [0, 0]]]))


def test_geo_dropin_classes_match_reference(golden):
    """AstarPlanner.plan / GeoPlanner.prune_path_nodes / geo_traj_plan (GEO:19-39) with the reference's RNG stream."""
    g = golden('geo_M3.npz')
    off = np.concatenate(([0], np.cumsum(g['path_len'])))
    e = ESDF()
    e.occupancy_map_cb(make_world(0).occupancy_msg())
    cfg = YamlConfig()
    ap = AstarPlanner()
    first = np.nonzero((g['world_id'] == 0) & (g['dense'] == 0))[0]
    for i in first[:6]:
        path = ap.plan(e, g['head'][i, 0], g['tail'][i, 0])
        assert isinstance(path, list) and np.array_equal(np.array(path), g['path'][off[i]:off[i + 1]])
    gp = GeoPlanner(cfg)
    i = int(first[np.argmax(g['n_keys'][first])])
    assert np.array_equal(np.array(gp.prune_path_nodes(e, g['path'][off[i]:off[i + 1]].tolist())), g['pruned'][i])
    agree = n = 0
    for k, i in enumerate(first[:24]):
        gp = GeoPlanner(cfg)
        st = NS(global_pos=np.array([g['head'][i, 0, 0], g['head'][i, 0, 1], 2.0]),
                global_vel=np.array([g['head'][i, 1, 0], g['head'][i, 1, 1], 0.0]))
        np.random.seed(100 + k)
        try:
            gp.geo_traj_plan(e, st, g['tail'][i]); ok = 1
        except Exception as ex:
            ok = 0
            assert 'No solution' in str(ex)
        n += 1
        assert np.array_equal(np.array(gp.path_pruned), g['pruned'][i])
        if ok != g['plan_ok'][i]:
            continue
        if ok:
            x = np.concatenate((gp.int_wpts.reshape(-1), gp.tau))
            agree += int(np.max(np.abs(x - g['plan_x'][i])) < 1e-6 and gp.iter_num == g['plan_iter'][i]
                         and gp.opt_running_times == g['plan_runs'][i])
        else:
            agree += int(gp.opt_running_times == g['plan_runs'][i])
    print(f'GeoPlanner.geo_traj_plan: {agree}/{n} identical to the reference')
    assert agree >= n - 2


def test_batch_geo_plan_equals_single_calls():
    """BatchGeoPlanner.geo_plan = k_astar + k_optimize for B problems; same results as warm-starting each problem from the
    oracle's pruned waypoints, and a usable success rate."""
    w = make_world(3)
    bp = BatchGeoPlanner(YamlConfig(), max_maps=1)
    bp.set_map(w)
    gm = _grid(w)
    head, tail = make_problems(w, 256, M=3)
    res = bp.geo_plan(head, tail, rng=np.random.default_rng(5))
    assert np.all(res['astar_status'] == lib.ASTAR_FOUND) and res['ok'].mean() > 0.75
    for i in range(0, 256, 16):
        iw, ts, _ = astar_ref.geo_guess(gm, head[i, 0], tail[i, 0], 2.5)
        assert np.array_equal(iw, res['pruned'][i, 1:3].T)
    one = bp.warm_start_plan(head[:32], tail[:32], res['pruned'][:32, 1:3].transpose(0, 2, 1), np.tile([3.75, 2.5, 3.75], (32, 1)),
                             max_attempts=1)
    first = res['attempt'][:32] == 0
    assert first.sum() > 12 and np.array_equal(one['x'][first], res['x'][:32][first])


def test_astar_second_pass_gives_the_same_results(monkeypatch):
    """k_astar's first pass has room for ASTAR_INSERT_CAP inserted nodes per search; searches that outgrow it are re-run
    by a second pass with full-size lists. With a first pass of 64 / 700 entries (NEO_ASTAR_ICAP, read at neo_create)
    most / some searches take the second pass: statuses, paths, key nodes and expansion counts must not change."""
    w = make_world(11)
    head, tail = make_problems(w, 160, M=6)
    outs = []
    for icap in (None, 64, 700):
        if icap is None:
            monkeypatch.delenv('NEO_ASTAR_ICAP', raising=False)
        else:
            monkeypatch.setenv('NEO_ASTAR_ICAP', str(icap))
        bp = BatchGeoPlanner(YamlConfig(), max_maps=1)
        bp.set_map(w)
        outs.append(bp.handle.astar(head[:, 0], tail[:, 0], max_path=1024))
        outs.append(bp.handle.astar(head[:, 0], tail[:, 0], max_path=1024))          # scratch left clean by both passes
    assert outs[0]['closed'].max() > 700                                             # the forced caps are exceeded
    for o in outs[1:]:
        for k in ('status', 'path_len', 'closed', 'pruned', 'path'):
            assert np.array_equal(o[k], outs[0][k]), k

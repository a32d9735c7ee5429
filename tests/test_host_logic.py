"""CPU tests of the host-side logic: initial guesses (EP:82-140) are bit-identical to the Python restatement of the
reference (oracle/minco_ref.py, itself pinned to the reference by gen_golden.py), and the C-ABI library loads
and exports every symbol include/neoopt.h declares (no compute calls without a GPU)."""
import os
import re

import numpy as np
import pytest

from neo_planner_b200 import guesses, lib
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig
from oracle import minco_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('M', [2, 3, 6, 10])
def test_straight_line_guess_bit_identical(M):
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    w = make_world(2)
    head, tail = make_problems(w, 64, M=M)
    head[5, 0, 1] = tail[5, 0, 1] = 1.25          # motion exactly along x: numpy.linspace's step == 0 branch
    head[6, 0, 0] = tail[6, 0, 0] = 3.0
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    ref = minco_ref.RefOptimizer(cfg)
    for b in range(64):
        q, ts = ref.straight_line_guess(head[b], tail[b])
        assert np.array_equal(q0[b], q) and np.array_equal(ts0[b], ts), b


def test_retry_guesses_follow_global_rng_stream():
    cfg = YamlConfig(); M = 3
    w = make_world(2)
    head, tail = make_problems(w, 4)
    ref = minco_ref.RefOptimizer(cfg)
    for b in range(4):
        np.random.seed(42 + b)
        mine, rts = guesses.retry_guesses(cfg, head[b], tail[b], M, 4)
        np.random.seed(42 + b)
        for a in range(4):
            q, ts = ref.straight_line_guess(head[b], tail[b], seed=a + 1)
            assert np.array_equal(mine[a], q) and np.array_equal(rts, ts)


def test_lateral_guesses_bit_identical():
    cfg = YamlConfig(); M = 3
    w = make_world(2)
    head, tail = make_problems(w, 32)
    cands, ts = guesses.lateral_guesses(cfg, head, tail, M)
    ref = minco_ref.RefOptimizer(cfg)
    for b in range(32):
        c, t = ref.lateral_guesses(head[b], tail[b])
        assert np.array_equal(cands[b], c) and np.array_equal(ts, t)


def test_adaptive_piece_count():
    cfg = YamlConfig(); cfg.init_wpts_mode = 'adaptive'
    head = np.array([[0.0, 0.0], [0, 0]]); tail = np.array([[5.0, 0.0], [0, 0]])
    assert guesses.pieces_for(cfg, head, tail) == 3          # ceil(5/2 - 1) = 2 waypoints
    tail[0, 0] = 1.0
    assert guesses.pieces_for(cfg, head, tail) == 2          # max(.., 1) waypoint


def test_library_exports_every_declared_symbol():
    from neo_planner_b200 import build
    build.build()
    hdr = open(os.path.join(ROOT, 'include', 'neoopt.h')).read()
    declared = set(re.findall(r'^(?:int|const char \*)\s*\*?(neo_\w+)\(', hdr, flags=re.M))
    assert len(declared) >= 20
    l = lib.load()
    for name in declared:
        assert hasattr(l, name), name
    assert declared == set(lib.EXPORTS)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(lib.NeoError, match='no usable CUDA device'):
        lib.Handle(YamlConfig())


def test_exp_dd_host_build_is_correctly_rounded():
    """The double-double exp (csrc/dd_exp.h), compiled for the host from the same header the kernels use, against
    mpmath at 200 bits: correctly rounded on every sample; and equal to libm's exp at the structural points
    tau0 = map_T2tau(k * 0.1) where a 1-ulp difference would flip int(T/delta_t) (EP:401)."""
    import math
    import mpmath as mp
    mp.mp.prec = 200
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-20, 20, 3000), rng.uniform(-700, 700, 500)])
    y = lib.exp_host(x)
    for xi, yi in zip(x, y):
        assert float(mp.exp(mp.mpf(float(xi)))) == yi, xi
    for T_min, T_max in [(0.5, 5.0), (2.0, 20.0)]:
        ts = np.arange(int(T_min * 10) + 1, int(T_max * 10)) * 0.1
        for T in np.concatenate([ts, ts * 1.5]):
            if not (T_min < T < T_max):
                continue
            tau = -math.log((T_max - T_min) / (T - T_min) - 1)
            e1 = lib.exp_host(np.array([-tau]))[0]
            T1 = (T_max - T_min) / (1 + e1) + T_min
            T2 = (T_max - T_min) / (1 + math.exp(-tau)) + T_min
            assert T1 == T2, T


def test_frames_against_scipy_rotation():
    from scipy.spatial.transform import Rotation
    from neo_planner_b200 import frames
    rng = np.random.default_rng(3)
    q = rng.normal(size=(50, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)      # (w, x, y, z)
    v = rng.normal(size=(50, 3))
    rot = Rotation.from_quat(q[:, [1, 2, 3, 0]])                                      # scipy wants (x, y, z, w)
    assert np.allclose(frames.rotation_matrix(q), rot.as_matrix(), atol=1e-14)
    assert np.allclose(frames.rotate(q, v), rot.apply(v), atol=1e-13)
    assert np.allclose(frames.rotate_inverse(q, v), rot.inv().apply(v), atol=1e-13)


def test_nn_io_layout_roundtrip():
    """form_nn_output (record_planner.py:61-72) and get_wpts_world (nn_planner.py:123-134) are inverse maps; the motion
    vector has the reference's 24-float layout (record_planner.py:43-48)."""
    from neo_planner_b200 import frames
    rng = np.random.default_rng(4)
    B = 6
    q = rng.normal(size=(B, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    gp = rng.normal(size=(B, 3)); gv = rng.normal(size=(B, 3)); lv = rng.normal(size=(B, 3))
    iw = rng.normal(size=(B, 2, 2)) + 3
    local = frames.form_nn_output(q, gp, 2.0, iw)
    back, ts = frames.wpts_world(q, gp, np.concatenate([local, np.ones((B, 3))], axis=1))
    assert np.allclose(back, iw, atol=1e-12) and ts.shape == (B, 3)
    depth = rng.uniform(0.2, 9.0, size=(B, 48, 64))
    target = rng.normal(size=(B, 2, 2))
    dn, motion = frames.form_nn_input(depth, lv, q, gp, gv, 2.0, gp + 0.1, gv, target)
    assert dn.dtype == np.uint8 and dn.max() == 255 and motion.shape == (B, 24)
    assert np.allclose(motion[:, :3], lv) and np.allclose(motion[:, 3:12], frames.rotation_matrix(q).reshape(B, 9))
    tp = np.concatenate([target[:, 0], np.full((B, 1), 2.0)], axis=1)
    assert np.allclose(frames.rotate(q, motion[:, 18:21]) + gp, tp, atol=1e-12)


def test_record_schema_readable_by_reference_trainer(tmp_path):
    """Rows follow record_planner.py:95-129; the reference's DataReader slices row[1:-9] / row[-9:] (nn_trainer.py:71-94)."""
    import pandas as pd
    from neo_planner_b200 import record
    assert record.TABLE_HEADER[:4] == ['id', 'drone_vel_x', 'drone_vel_y', 'drone_vel_z']
    assert record.TABLE_HEADER[4:13] == ['R11', 'R12', 'R13', 'R21', 'R22', 'R23', 'R31', 'R32', 'R33']
    assert record.TABLE_HEADER[13:25] == ['init_pos_x', 'init_pos_y', 'init_pos_z', 'init_vel_x', 'init_vel_y', 'init_vel_z',
                                          'target_pos_x', 'target_pos_y', 'target_pos_z', 'target_vel_x', 'target_vel_y',
                                          'target_vel_z']
    assert record.TABLE_HEADER[25:] == ['wpts1_x', 'wpts1_y', 'wpts1_z', 'wpts2_x', 'wpts2_y', 'wpts2_z', 'ts1', 'ts2', 'ts3']
    rng = np.random.default_rng(0)
    motion = rng.normal(size=(5, 24)); wl = rng.normal(size=(5, 6)); ts = rng.uniform(1, 4, size=(5, 3))
    ids = record.make_ids(5)
    csv = tmp_path / 'training_data' / 'train.csv'
    record.append_csv(str(csv), record.training_rows(ids[:3], motion[:3], wl[:3], ts[:3]))
    record.append_csv(str(csv), record.training_rows(ids[3:], motion[3:], wl[3:], ts[3:]))
    back = pd.read_csv(csv)
    assert list(back.columns) == record.TABLE_HEADER and len(back) == 5 and len(set(back['id'])) == 5
    for i, row in back.iterrows():
        assert row['id'][0] == 't' and int(row['id'][1:]) > 0
        assert np.allclose(row.iloc[1:-9].values.astype(float), motion[i]) and np.allclose(row.iloc[-9:].values.astype(float),
                                                                                          np.concatenate([wl[i], ts[i]]))


def test_initializer_network_shapes():
    import torch
    from neo_planner_b200 import initializer
    torch.manual_seed(0)
    for cls in (initializer.PlannerNet, initializer.PlannerNetConv):
        net = cls().eval()
        x = torch.rand(2, initializer.IMG_WIDTH * initializer.IMG_HEIGHT + 24)
        with torch.no_grad():
            y = net(x)
        assert y.shape == (2, 9) and torch.isfinite(y).all()
    n_params = sum(p.numel() for p in initializer.PlannerNet().parameters())
    assert 11_000_000 < n_params < 12_000_000          # ResNet-18 trunk (1-channel stem) + the small heads


def test_map_cache_reuploads_a_new_map_object():
    """The upload cache is keyed on the map object's identity; a map that is garbage-collected must not hand its id() to
    the next one unnoticed (the cache holds a reference to what it cached)."""
    from types import SimpleNamespace as NS
    from neo_planner_b200.planner import _MapCache

    class FakeHandle:
        def __init__(self):
            self.uploads = []

        def set_map_occupancy(self, slot, H, W, res, ox, oy, occ):
            self.uploads.append((slot, int(np.asarray(occ).sum())))

    fh = FakeHandle()
    cache = _MapCache(fh)
    for k in range(6):                       # temporaries: each is freed before the next is created
        cache.ensure(0, NS(occ=np.full((4, 4), k, np.int8), H=4, W=4, res=1.0, ox=0.0, oy=0.0))
    assert [u[1] for u in fh.uploads] == [16 * k for k in range(6)]
    m = NS(occ=np.zeros((4, 4), np.int8), H=4, W=4, res=1.0, ox=0.0, oy=0.0)
    cache.ensure(0, m); cache.ensure(0, m)
    assert len(fh.uploads) == 7


def test_four_key_indices_follow_the_reference_rule():
    """geo.four_key_indices (GEO:78-95) against the checker's statement of the same rule on random key lists."""
    from neo_planner_b200.geo import four_key_indices, geo_times
    from oracle import astar_ref
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 4, 5, 6, 9, 17):
        for _ in range(40):
            keys = [0] if n == 1 else [0] + sorted(rng.choice(np.arange(1, 400), size=n - 1, replace=False).tolist())
            got = four_key_indices(keys)
            assert got == astar_ref.four_of(keys) and len(got) == 4 and got[0] == 0 and got[-1] == keys[-1]
            assert all(isinstance(v, int) for v in got)
    ts = geo_times(YamlConfig())
    assert ts.tolist() == [3.75, 2.5, 3.75] and geo_times(YamlConfig(), 5).shape == (5, 3)


def test_T2tau_host_matches_python_math_with_repeats():
    """neo_T2tau (host libm log behind a small per-call memo) against the reference's formula (EP:468-475) evaluated with
    Python's math.log: bit-identical for repeated, distinct and invalid durations in one call."""
    import ctypes as C
    import math
    cfg = lib.Config.from_config(YamlConfig())
    rng = np.random.default_rng(3)
    base = np.array([3.75, 2.5, 3.75, 0.5, 5.0, 0.4, 5.5, 0.5000001, 4.9999999])
    ts = np.concatenate([base, rng.uniform(0.5, 5.0, 5000), np.tile(base, 40), rng.choice(rng.uniform(0.6, 4.9, 37), 3000)])
    tau = np.zeros_like(ts); st = np.zeros(ts.size, np.int32)
    assert lib.load().neo_T2tau(C.byref(cfg), ts.size, lib.ptr(ts), lib.ptr(tau), lib.ptr(st)) == 0
    for T, t, s in zip(ts, tau, st):
        try:
            want = -math.log((5.0 - 0.5) / (float(T) - 0.5) - 1)      # Python floats: T == T_min divides by zero
            assert s == 0 and t == want, (T, t, want)
        except (ValueError, ZeroDivisionError):
            assert s == lib.ST_DOMAIN and t == 0.0, (T, s)
    bad = np.array([np.nan, 2.5]); tau = np.zeros(2); st = np.zeros(2, np.int32)      # NaN durations are rejected up front
    lib.load().neo_T2tau(C.byref(cfg), 2, lib.ptr(bad), lib.ptr(tau), lib.ptr(st))
    assert st.tolist() == [lib.ST_DOMAIN, 0]


# ---- pins generated from the unmodified reference (oracle/gen_golden.py nn_io nets) --------------------------------
def test_nn_io_against_reference_functions(golden):
    """frames.py vs the reference's own form_nn_input / form_nn_output (record_planner.py:13-72) and
    NNPlanner.get_wpts_world (nn_planner.py:123-134), run unmodified on seeded drone states when the fixture was made
    (with a pyquaternion stand-in: q v q* through the 4x4 product matrices, like pyquaternion 0.9)."""
    from neo_planner_b200 import frames
    g = golden('nn_io.npz')
    dn, mi = frames.form_nn_input(g['depth'], g['local_vel'], g['quat'], g['global_pos'], g['global_vel'], float(g['des_pos_z']),
                                  g['init_pos'], g['init_vel'], g['target'])
    assert dn.dtype == np.uint8 and np.array_equal(dn, g['depth_norm'])                # truncation to uint8: exact
    assert mi.shape == g['motion'].shape == (12, 24)
    assert np.max(np.abs(mi - g['motion'])) <= 1e-13 * max(1.0, np.max(np.abs(g['motion'])))
    loc = frames.form_nn_output(g['quat'], g['global_pos'], float(g['des_pos_z']), g['int_wpts'])
    assert np.max(np.abs(loc - g['int_wpts_local'])) <= 1e-13 * np.max(np.abs(g['int_wpts_local']))
    # network output layout (nn_planner.py:104-105): [wpt1 xyz, wpt2 xyz, ts1..3] -> (3, 2) body-frame columns
    net_out = np.concatenate([np.transpose(g['net_local'], (0, 2, 1)).reshape(12, 6), np.ones((12, 3))], axis=1)
    world, ts = frames.wpts_world(g['quat'], g['global_pos'], net_out)
    assert np.max(np.abs(world - g['wpts_world'][:, :2, :])) <= 1e-13 * np.max(np.abs(g['wpts_world']))
    assert np.array_equal(ts, np.ones((12, 3)))


@pytest.mark.parametrize('tag', ['mlp', 'conv'])
def test_initializer_networks_match_the_reference_classes(golden, tag):
    """Structural parity of initializer.PlannerNet / PlannerNetConv with nn_trainer.py:109-155 / nn_trainer_conv.py:108-160:
    identical state_dict keys and shapes, and -- with the same deterministic weights (oracle/net_weights.py) on the same
    seeded input -- the fp32 outputs the reference classes produced when the fixture was generated."""
    import torch
    from neo_planner_b200 import initializer
    from oracle import net_weights
    g = golden('nets.npz')
    net = (initializer.PlannerNet() if tag == 'mlp' else initializer.PlannerNetConv()).eval()
    assert net_weights.signature(net) == [str(s) for s in g[tag + '_sig']]
    net_weights.fill_deterministic(net)
    with torch.no_grad():
        y = net(torch.from_numpy(net_weights.sample_input())).reshape(2, -1).numpy()
    ref = g[tag + '_out']
    assert y.shape == ref.shape == (2, 9)
    assert np.max(np.abs(y - ref)) <= 1e-4 * max(1.0, np.max(np.abs(ref))), np.max(np.abs(y - ref))


def test_record_ids_are_unique_across_calls_and_ranks():
    """two batches recorded within the same few milliseconds must not share ids (the id names the depth PNG)."""
    import datetime
    from neo_planner_b200 import record
    now = datetime.datetime(2026, 1, 2, 3, 4, 5, 678000)
    a = record.make_ids(1024, now=now)
    b = record.make_ids(1024, now=now + datetime.timedelta(milliseconds=100))
    assert len(set(a)) == 1024 and not set(a) & set(b)
    assert all(i[0] == 't' and i[1:].isdigit() for i in a)
    r0 = record.make_ids(64, now=now, rank=0, world_size=4)
    r3 = record.make_ids(64, now=now, rank=3, world_size=4)
    assert not set(r0) & set(r3) and not (set(r0) | set(r3)) & (set(a) | set(b))


def test_clamp_durations_flags_out_of_range_predictions():
    from neo_planner_b200 import frames
    ts = np.array([[1.0, 2.0, 3.0], [0.2, 2.0, 9.0], [np.nan, 1.0, 1.0]])
    out, touched = frames.clamp_durations(ts, 0.5, 5.0)
    assert touched.tolist() == [False, True, True] and np.array_equal(out[0], ts[0])
    assert ((out > 0.5) & (out < 5.0)).all()


def test_host_worker_pool_covers_every_index_once():
    """neo_optimize assembles inputs and scatters results of a large batch on a persistent pool of host threads
    (csrc/neoopt.cu: HostPool). The test hook runs the pool without a GPU: over many calls and sizes -- below the
    threading threshold, not divisible by the thread count, large -- every index is visited exactly once per call."""
    import ctypes
    l = lib.load()
    l.neo_test_host_pool.argtypes = [ctypes.c_longlong, ctypes.c_int]
    l.neo_test_host_pool.restype = ctypes.c_int
    for count, calls in ((1, 3), (4095, 3), (8192, 50), (65536, 200), (65537, 50), (100003, 50), (1 << 20, 10)):
        assert l.neo_test_host_pool(count, calls) == 0, count

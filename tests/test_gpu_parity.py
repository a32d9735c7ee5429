"""GPU parity tests: the sm_100a path (through the C ABI, libneoopt.so) against
  * golden vectors produced by the unmodified reference (tests/golden, oracle/gen_golden.py), and
  * the plain-C CPU checker (oracle/minco_oracle.c) on seeded inputs.
Bit-exact for map build / cell indexing; <= 1e-6 relative for cost and gradient; <= 1e-4 m for final
trajectory coefficients (BASELINE.json north_star)."""
import numpy as np
import pytest

from neo_planner_b200 import lib
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig, LibraryDefaultConfig
from neo_planner_b200.guesses import straight_line_guess, retry_guesses
from oracle import c_oracle

pytestmark = pytest.mark.gpu


class Cfg:
    def __init__(self, v):
        (self.v_max, self.T_min, self.T_max, self.safe_dis, self.delta_t) = [float(t) for t in v[:5]]
        self.weights = [float(t) for t in v[5:9]]
        self.collision_cost_tol = float(v[9])
        self.init_T = float(v[10])
        self.init_wpts_mode = 'fixed'; self.init_seg_len = 2.0; self.init_wpts_num = 2; self.opt_tol = 1e-4


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope='module')
def world0():
    return make_world(0)


def handle_for(cfg, world, max_maps=1):
    h = lib.Handle(cfg, 0, max_maps)
    h.set_map_occupancy(0, world.H, world.W, world.res, world.ox, world.oy, world.occ)
    return h


# ------------------------------------------------------------------------------------------ exp
def test_exp_device_matches_host_build():
    h = lib.Handle(YamlConfig())
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-30, 30, 400000), rng.uniform(-745, 709.78, 100000),
                        -np.log(4.5 / (np.arange(6, 50) * 0.1 - 0.5) - 1 + 1e-300)])
    assert np.array_equal(h.exp_dev(x), lib.exp_host(x))


# ------------------------------------------------------------------------------------------ map
def test_map_build_bit_exact_golden(golden):
    g = golden('esdf_small.npz')
    H, W = int(g['H']), int(g['W'])
    h = lib.Handle(YamlConfig())
    h.set_map_occupancy(0, H, W, float(g['res']), float(g['ox']), float(g['oy']), g['occ'])
    e, gx, gy = h.get_map(0, H, W)
    assert np.array_equal(e, g['esdf']) and np.array_equal(gx, g['gx']) and np.array_equal(gy, g['gy'])
    idx, d, gr = h.query_map(0, g['pts'])
    assert np.array_equal(idx, g['idx']) and np.array_equal(d, g['dis']) and np.array_equal(gr, g['grad'])
    # upload path (arrays from an existing ESDF object) gives the same lookups
    h.set_map_esdf(0, float(g['res']), float(g['ox']), float(g['oy']), g['esdf'], g['gx'], g['gy'])
    idx2, d2, gr2 = h.query_map(0, g['pts'])
    assert np.array_equal(idx2, g['idx']) and np.array_equal(d2, g['dis']) and np.array_equal(gr2, g['grad'])
    # all-free map: scipy's virtual feature at (-1, 0)
    h.set_map_occupancy(0, 9, 11, 0.2, 0.0, 0.0, np.zeros((9, 11), np.int8))
    e, gx, gy = h.get_map(0, 9, 11)
    assert np.array_equal(e, g['free_esdf']) and np.array_equal(gx, g['free_gx']) and np.array_equal(gy, g['free_gy'])


@pytest.mark.parametrize('dense', [False, True])
def test_map_build_bit_exact_worlds(dense):
    w = make_world(3, dense=dense)
    h = handle_for(YamlConfig(), w)
    e, gx, gy = h.get_map(0, w.H, w.W)
    m = c_oracle.OracleMap.from_world(w)
    assert np.array_equal(e, m.esdf) and np.array_equal(gx, m.gx) and np.array_equal(gy, m.gy)
    rng = np.random.default_rng(1)
    pts = rng.uniform([w.ox - 1, w.oy - 1], [w.ox + w.W * w.res + 1, w.oy + w.H * w.res + 1], size=(20000, 2))
    a = h.query_map(0, pts); b = m.query(pts)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


def test_map_build_random_small():
    rng = np.random.default_rng(9)
    h = lib.Handle(YamlConfig())
    for H, W, p in [(2, 2, 0.5), (2, 37, 0.1), (41, 3, 0.1), (64, 64, 0.001), (33, 65, 0.3)]:
        occ = np.where(rng.random((H, W)) < p, 100, 0).astype(np.int8)
        h.set_map_occupancy(0, H, W, 0.05, -1.0, 2.0, occ)
        e, gx, gy = h.get_map(0, H, W)
        m = c_oracle.OracleMap(occ, H, W, 0.05, -1.0, 2.0)
        assert np.array_equal(e, m.esdf) and np.array_equal(gx, m.gx) and np.array_equal(gy, m.gy), (H, W)


# ------------------------------------------------------------------------------------------ cost + gradient
@pytest.mark.parametrize('name', ['eval_M3.npz', 'eval_M10.npz', 'eval_M3_libdefaults.npz'])
def test_eval_against_reference_golden(golden, name):
    g = golden(name)
    M = int(g['M']); cfg = Cfg(g['cfg'])
    h = handle_for(cfg, make_world(int(g['world_id'])))
    out = h.eval(M, g['x'], g['head'], g['tail'], want_coeffs=True)
    assert (out['status'] == 0).all()
    w = np.array(cfg.weights)
    worst_c = worst_g = 0.0
    for k in range(len(g['x'])):
        f = out['costs'][k] @ w
        assert abs(f - g['f'][k]) <= 1e-6 * abs(g['f'][k]), (k, f, g['f'][k])
        assert np.allclose(out['costs'][k], g['costs'][k], rtol=1e-6, atol=1e-9)
        assert rel(out['grad'][k], g['grad'][k]) <= 1e-6, (k, rel(out['grad'][k], g['grad'][k]))
        assert np.max(np.abs(out['coeffs'][k] - g['coeffs'][k])) <= 1e-9
        worst_c = max(worst_c, abs(f - g['f'][k]) / abs(g['f'][k])); worst_g = max(worst_g, rel(out['grad'][k], g['grad'][k]))
    print(f'{name}: worst rel cost err {worst_c:.2e}, worst rel grad err {worst_g:.2e}')


@pytest.mark.parametrize('M', [2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_eval_against_c_oracle(M, world0):
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    B = 1024 if M == 3 else 192
    head, tail = make_problems(world0, B, M=M)
    q0, ts0 = straight_line_guess(cfg, head, tail, M)
    h = handle_for(cfg, world0)
    tau, st = h.T2tau(ts0)
    rng = np.random.default_rng(M)
    x = np.concatenate([q0.reshape(B, -1) + rng.normal(0, 0.4, (B, 2 * (M - 1))), tau + rng.normal(0, 0.8, (B, M))], axis=1)
    out = h.eval(M, x, head, tail)
    m = c_oracle.OracleMap.from_world(world0)
    costs, grad, status = c_oracle.eval_batch(c_oracle.Params.from_config(cfg), m, M, head, tail, x)
    assert np.array_equal(out['status'], status)
    w = np.array(cfg.weights, dtype=float)
    f_dev = out['costs'] @ w; f_ref = costs @ w
    assert np.max(np.abs(f_dev - f_ref) / np.abs(f_ref)) <= 1e-6
    gerr = np.max(np.abs(out['grad'] - grad), axis=1) / np.max(np.abs(grad), axis=1)
    assert gerr.max() <= 1e-6, gerr.max()
    assert (costs[:, 3] > 0).sum() > 0 and (costs[:, 2] > 0).sum() > 0     # both penalties exercised
    print(f'M={M}: worst rel cost err {np.max(np.abs(f_dev - f_ref) / np.abs(f_ref)):.2e}, worst rel grad err {gerr.max():.2e}')


def test_eval_overflow_status(world0):
    cfg = YamlConfig()
    head, tail = make_problems(world0, 4)
    q0, ts0 = straight_line_guess(cfg, head, tail, 3)
    h = handle_for(cfg, world0)
    tau, _ = h.T2tau(ts0)
    x = np.concatenate([q0.reshape(4, -1), tau], axis=1)
    x[1, 5] = -720.0      # math.exp(720) overflows (EP:481)
    x[2, 4] = -400.0      # (1+exp(400))**2 overflows (EP:490)
    x[3, 6] = np.nan
    out = h.eval(3, x, head, tail)
    m = c_oracle.OracleMap.from_world(world0)
    _, _, status = c_oracle.eval_batch(c_oracle.Params.from_config(cfg), m, 3, head, tail, x)
    assert list(out['status']) == [0, lib.ST_OVERFLOW, lib.ST_OVERFLOW, lib.ST_NAN]
    assert list(status[:3]) == [0, lib.ST_OVERFLOW, lib.ST_OVERFLOW]


# ------------------------------------------------------------------------------------------ optimisation
def _cls(msg):
    if msg.startswith('CONVERGENCE: REL'):
        return 0
    if msg.startswith('CONVERGENCE: NORM'):
        return 1
    if msg.startswith('ABNORMAL'):
        return 2
    return 4


@pytest.mark.parametrize('name,min_match', [('plans_M3.npz', 0.97), ('plans_M10.npz', 0.95)])
def test_plan_once_against_reference_golden(golden, name, min_match):
    """plan_once from the expert guess: same termination class and iteration count as scipy's L-BFGS-B in the
    reference, final decision vector within 1e-6 and coefficients within 1e-4 m, on all but the line-search
    knife-edge problems (every differing outcome is traced to a last-bit difference of f or g by
    tests/test_gpu_lockstep.py; the floor here is the measured rate minus two problems)."""
    g = golden(name)
    M = int(g['M']); cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    w = make_world(int(g['world_id']))
    h = handle_for(cfg, w)
    B = len(g['x0'])
    q0, ts0 = straight_line_guess(cfg, g['head'], g['tail'], M)
    assert np.array_equal(q0.reshape(B, -1), g['x0'][:, :2 * (M - 1)])
    out = h.optimize(M, q0, ts0, g['head'], g['tail'], max_attempts=1)
    match = 0
    for k in range(B):
        cls = _cls(str(g['msg'][k]))
        if cls == 4:
            match += int(out['status'][k] >= 4)
            continue
        same = (out['status'][k] == cls and out['nit'][k] == g['nit'][k] and np.max(np.abs(out['x'][k] - g['x'][k])) < 1e-6)
        match += int(same)
        if same:
            assert np.allclose(out['costs'][k], g['costs'][k], rtol=1e-6, atol=1e-9)
    print(f'{name}: {match}/{B} follow the reference optimizer')
    assert match >= min_match * B


def test_plan_with_retries_against_reference_golden(golden):
    """Full plan() (EP:62-80) with the retry noise the reference drew (np.random.seed(k), EP:94)."""
    g = golden('plans_M3.npz')
    cfg = YamlConfig(); M = 3
    h = handle_for(cfg, make_world(int(g['world_id'])))
    B = len(g['head'])
    q0, ts0 = straight_line_guess(cfg, g['head'], g['tail'], M)
    rq = np.zeros((B, 4, 2, M - 1))
    for k in range(B):
        np.random.seed(k)
        rq[k], rts = retry_guesses(cfg, g['head'][k], g['tail'][k], M, 4)
    out = h.optimize(M, q0, ts0, g['head'], g['tail'], retry_q=rq, retry_ts=rts, max_attempts=5)
    agree = 0
    for k in range(B):
        if out['ok'][k] != g['plan_ok'][k]:
            continue
        if not out['ok'][k]:
            agree += int(out['runs'][k] == g['plan_runs'][k])
            continue
        good = (np.max(np.abs(out['coeffs'][k] - g['plan_coeffs'][k])) <= 1e-4 and out['runs'][k] == g['plan_runs'][k]
                and out['nit'][k] == g['plan_iter'][k])
        agree += int(good)
    print(f'plan(): {agree}/{B} identical outcome (ok flag, runs, iterations, coefficients <= 1e-4 m)')
    assert agree >= 0.97 * B


@pytest.mark.parametrize('M,B', [(3, 1024), (10, 256)])
def test_optimize_against_c_oracle(M, B, world0):
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    head, tail = make_problems(world0, B, M=M)
    q0, ts0 = straight_line_guess(cfg, head, tail, M)
    rng = np.random.default_rng(11)
    rq, rts = retry_guesses(cfg, head, tail, M, 4, rng=rng)
    h = handle_for(cfg, world0)
    out = h.optimize(M, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=5)
    m = c_oracle.OracleMap.from_world(world0)
    ref = c_oracle.plan_batch(c_oracle.Params.from_config(cfg), m, M, head, tail, q0, ts0, rq, rts, 5)
    same_ok = (out['ok'] == ref['ok'])
    close = np.max(np.abs(out['coeffs'] - ref['coeffs']).reshape(B, -1), axis=1) <= 1e-4
    same_path = (out['runs'] == ref['runs']) & (out['nit'] == ref['nit']) & (out['status'] == ref['status'])
    good = same_ok & close & same_path
    print(f'M={M}: {good.sum()}/{B} identical to the CPU checker (ok {same_ok.mean():.3f}, coeffs<=1e-4 {close.mean():.3f}, '
          f'same path {same_path.mean():.3f}); ok rate {out["ok"].mean():.3f}; mean nfev {out["nfev"].mean():.1f}')
    assert good.mean() >= 0.975
    # size-independent properties on every problem: trajectory endpoints and continuity
    ok = out['ok'] == 1
    c = out['coeffs'][ok]; ts = out['ts'][ok]
    assert np.max(np.abs(c[:, 0, :] - head[ok][:, 0, :])) < 1e-9               # starts at the start
    T = ts[:, -1]
    pw = np.stack([T ** k for k in range(6)], axis=1)
    end = np.einsum('bk,bkd->bd', pw, c[:, -6:, :])
    assert np.max(np.abs(end - tail[ok][:, 0, :])) < 1e-7                       # ends at the goal
    assert ((ts > cfg.T_min) & (ts < cfg.T_max)).all()
    assert (out['costs'][ok][:, 3] * cfg.weights[3] <= cfg.collision_cost_tol).all()


def test_domain_error_behaviour(world0, golden):
    """Library defaults have init_T == T_min: every attempt dies in map_T2tau (EP:474) -> 'No solution'."""
    cfg = LibraryDefaultConfig()
    head, tail = make_problems(world0, 8)
    q0, ts0 = straight_line_guess(cfg, head, tail, 3)
    rq, rts = retry_guesses(cfg, head, tail, 3, 4, rng=np.random.default_rng(0))
    h = handle_for(cfg, world0)
    out = h.optimize(3, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=5)
    assert (out['ok'] == 0).all() and (out['runs'] == 0).all() and (out['status'] == lib.ST_DOMAIN).all()
    # a bad first guess only loses the first attempt (golden: errors.npz)
    g = golden('errors.npz')
    cfg = YamlConfig()
    h = handle_for(cfg, world0)
    q0, ts0 = straight_line_guess(cfg, g['head'][1:2], g['tail'][1:2], 3)
    np.random.seed(3)
    rq1, rts = retry_guesses(cfg, g['head'][1], g['tail'][1], 3, 4)
    out = h.optimize(3, q0, g['bad_ts'][None], g['head'][1:2], g['tail'][1:2], retry_q=rq1[None], retry_ts=rts, max_attempts=5)
    assert out['ok'][0] == 1 and out['runs'][0] == int(g['bad_runs']) and out['attempt'][0] == 1
    assert out['nit'][0] == int(g['bad_iter'])
    assert np.max(np.abs(out['x'][0] - g['bad_x'])) < 1e-6


# ------------------------------------------------------------------------------------------ coefficients / sampling
def test_get_coeffs_and_sampling(golden, world0):
    g = golden('plans_M3.npz')
    cfg = YamlConfig()
    h = handle_for(cfg, world0)
    ok = g['plan_ok'] == 1
    q = g['plan_x'][ok][:, :4].reshape(-1, 2, 2); ts = g['plan_ts'][ok]
    c = h.get_coeffs(3, q, ts, g['head'][ok], g['tail'][ok])
    assert np.max(np.abs(c - g['plan_coeffs'][ok])) < 1e-10
    k = int(g['cmd_index'])
    states, count = h.sample(3, g['plan_coeffs'][k:k + 1], g['plan_ts'][k:k + 1], 60.0)
    assert count[0] == g['cmd'].shape[0]
    assert np.max(np.abs(states[0, :count[0]] - g['cmd'])) < 1e-10
    # batch vs the C checker, ragged lengths
    states, count = h.sample(3, g['plan_coeffs'][ok], ts, 60.0)
    for i in range(0, ok.sum(), 7):
        ref = c_oracle.sample(3, g['plan_coeffs'][ok][i], ts[i], 60.0)
        assert count[i] == ref.shape[0] and np.max(np.abs(states[i, :count[i]] - ref)) < 1e-10


def test_multi_map_slots():
    cfg = YamlConfig()
    worlds = [make_world(i) for i in (4, 5, 6)]
    h = lib.Handle(cfg, 0, 3)
    heads, tails, ids = [], [], []
    for s, w in enumerate(worlds):
        h.set_map_occupancy(s, w.H, w.W, w.res, w.ox, w.oy, w.occ)
        a, b = make_problems(w, 64)
        heads.append(a); tails.append(b); ids.append(np.full(64, s, np.int32))
    head = np.concatenate(heads); tail = np.concatenate(tails); ids = np.concatenate(ids)
    q0, ts0 = straight_line_guess(cfg, head, tail, 3)
    out = h.optimize(3, q0, ts0, head, tail, map_ids=ids, max_attempts=1)
    for s, w in enumerate(worlds):
        h1 = handle_for(cfg, w)
        sl = slice(64 * s, 64 * (s + 1))
        o1 = h1.optimize(3, q0[sl], ts0[sl], head[sl], tail[sl], max_attempts=1)
        assert np.array_equal(o1['x'], out['x'][sl]) and np.array_equal(o1['status'], out['status'][sl])


def test_plan_against_scipy_on_unseen_problems():
    """Problems that are in no fixture: the device path against oracle/minco_ref.py, which runs the REAL scipy L-BFGS-B
    on the reference's arithmetic (bit-identical to the reference where the fixtures were generated)."""
    from oracle import minco_ref
    cfg = YamlConfig(); M = 3
    w = make_world(21)
    B = 48
    head, tail = make_problems(w, B, first=500)
    q0, ts0 = straight_line_guess(cfg, head, tail, M)
    rq = np.zeros((B, 4, 2, M - 1))
    for k in range(B):
        np.random.seed(9000 + k)
        rq[k], rts = retry_guesses(cfg, head[k], tail[k], M, 4)
    h = handle_for(cfg, w)
    out = h.optimize(M, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=5)
    grid = minco_ref.GridMap(w.occ, w.H, w.W, w.res, w.ox, w.oy)
    agree = 0
    for k in range(B):
        opt = minco_ref.RefOptimizer(cfg)
        np.random.seed(9000 + k)
        try:
            opt.plan(grid, head[k], tail[k]); ok = 1
        except Exception:
            ok = 0
        if ok != out['ok'][k]:
            continue
        if not ok:
            agree += int(opt.opt_running_times == out['runs'][k])
            continue
        c = opt.final_coeffs()
        agree += int(np.max(np.abs(c - out['coeffs'][k])) <= 1e-4 and opt.iter_num == out['nit'][k]
                     and opt.opt_running_times == out['runs'][k])
    print(f'unseen problems vs scipy: {agree}/{B} identical (ok flag, attempts, iterations, coefficients <= 1e-4 m)')
    assert agree >= 0.97 * B


@pytest.mark.parametrize('M', [3, 6, 10])
def test_eval_extreme_time_allocations(M, world0):
    """Durations pushed against both bounds (T_i within 1e-3 of T_min or T_max, ratios up to 10 between neighbours): the
    un-pivoted block elimination of the node-state system must still agree with the dense pivoted solve of the checker."""
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    B = 256
    head, tail = make_problems(world0, B, M=M)
    q0, ts0 = straight_line_guess(cfg, head, tail, M)
    rng = np.random.default_rng(50 + M)
    tau = rng.choice([-8.0, -3.0, 0.0, 3.0, 8.0], size=(B, M)) + rng.normal(0, 0.2, (B, M))
    x = np.concatenate([q0.reshape(B, -1) + rng.normal(0, 0.3, (B, 2 * (M - 1))), tau], axis=1)
    h = handle_for(cfg, world0)
    out = h.eval(M, x, head, tail, want_coeffs=True)
    m = c_oracle.OracleMap.from_world(world0)
    costs, grad, status = c_oracle.eval_batch(c_oracle.Params.from_config(cfg), m, M, head, tail, x)
    assert (status == 0).all() and (out['status'] == 0).all()
    assert out['ts'].min() < 0.502 and out['ts'].max() > 4.998
    w = np.array(cfg.weights, dtype=float)
    f_dev = out['costs'] @ w; f_ref = costs @ w
    assert np.max(np.abs(f_dev - f_ref) / np.abs(f_ref)) <= 1e-6
    gerr = np.max(np.abs(out['grad'] - grad), axis=1) / np.max(np.abs(grad), axis=1)
    print(f'M={M}: extreme T: worst rel cost err {np.max(np.abs(f_dev - f_ref) / np.abs(f_ref)):.2e}, worst rel grad err {gerr.max():.2e}')
    assert gerr.max() <= 1e-6


def test_batched_map_build_equals_single_builds():
    """neo_set_maps_occupancy (K maps, one synchronisation) builds exactly what K neo_set_map_occupancy calls build, and
    both equal the CPU checker's scipy-exact EDT (ESDF:23-33)."""
    import time
    K = 12
    worlds = [make_world(40 + k) for k in range(K)]
    w0 = worlds[0]
    h1 = lib.Handle(YamlConfig(), 0, K); h2 = lib.Handle(YamlConfig(), 0, K)
    t0 = time.perf_counter()
    for k, w in enumerate(worlds):
        h1.set_map_occupancy(k, w.H, w.W, w.res, w.ox, w.oy, w.occ)
    t1 = time.perf_counter()
    h2.set_maps_occupancy(np.arange(K)[::-1].copy(), w0.H, w0.W, w0.res, [w.ox for w in worlds][::-1], [w.oy for w in worlds][::-1],
                          np.stack([np.asarray(w.occ).reshape(w.H, w.W) for w in worlds])[::-1].copy())
    t2 = time.perf_counter()
    for k, w in enumerate(worlds):
        a = h1.get_map(k, w.H, w.W); b = h2.get_map(k, w.H, w.W)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        m = c_oracle.OracleMap.from_world(w)
        assert np.array_equal(b[0], m.esdf) and np.array_equal(b[1], m.gx) and np.array_equal(b[2], m.gy)
    print(f'{K} maps: {1e3 * (t1 - t0):.2f} ms one by one, {1e3 * (t2 - t1):.2f} ms in one call')


def test_invalid_map_ids_are_rejected(world0):
    """a map id that names no uploaded map must be an API error, not an out-of-bounds read or a silently
    obstacle-free plan."""
    cfg = YamlConfig()
    head, tail = make_problems(world0, 4)
    q0, ts0 = straight_line_guess(cfg, head, tail, 3)
    h = lib.Handle(cfg, 0, 4)
    h.set_map_occupancy(1, world0.H, world0.W, world0.res, world0.ox, world0.oy, world0.occ)
    for ids in (np.array([1, 1, 7, 1], np.int32), np.array([1, -1, 1, 1], np.int32), np.array([1, 2, 1, 1], np.int32), None):
        with pytest.raises(lib.NeoError, match='map'):
            h.optimize(3, q0, ts0, head, tail, ids, max_attempts=1)
        tau, _ = h.T2tau(ts0)
        with pytest.raises(lib.NeoError, match='map'):
            h.eval(3, np.concatenate([q0.reshape(4, -1), tau], axis=1), head, tail, ids)
    out = h.optimize(3, q0, ts0, head, tail, np.array([1, 1, 1, 1], np.int32), max_attempts=1)      # the valid slot works
    assert out['status'].max() <= 6

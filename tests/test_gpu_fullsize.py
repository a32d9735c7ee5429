"""Full-size runs (BASELINE.json configs 4 and 5): every problem against the CPU checker (all host threads), plus the
size-independent properties -- determinism, independence of batch composition and order, agreement of the speculative
retry scheduler with plain sequential semantics, and the invariants every MINCO trajectory must satisfy (boundary
states, C^4 continuity at the waypoints, durations inside (T_min, T_max), collision tolerance)."""
import os

import numpy as np
import pytest

from neo_planner_b200 import guesses, lib
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def pinned_handle(cfg, n_maps, tile):
    """A handle whose lanes-per-problem choice does not depend on the batch size (NEO_TILE is read at neo_create)."""
    old = os.environ.get('NEO_TILE')
    os.environ['NEO_TILE'] = str(tile)
    try:
        return lib.Handle(cfg, 0, n_maps)
    finally:
        if old is None:
            del os.environ['NEO_TILE']
        else:
            os.environ['NEO_TILE'] = old


def compare_with_checker(cfg, M, worlds, ids, head, tail, q0, ts0, rq, rts, out, sel, floor):
    """Device results `out` (all problems) vs the multithreaded C checker on the problems `sel`."""
    maps = [c_oracle.OracleMap.from_world(w) for w in worlds]
    p = c_oracle.Params.from_config(cfg)
    ref = c_oracle.plan_batch_mt(p, maps, M, head[sel], tail[sel], q0[sel], ts0[sel], rq[sel], rts, 5,
                                 map_ids=None if ids is None else ids[sel])
    n = len(ref['ok'])
    same_path = ((out['ok'][sel] == ref['ok']) & (out['runs'][sel] == ref['runs']) & (out['nit'][sel] == ref['nit'])
                 & (out['status'][sel] == ref['status']) & (out['attempt'][sel] == ref['attempt']))
    dc = np.max(np.abs(out['coeffs'][sel] - ref['coeffs']).reshape(n, -1), axis=1)
    good = same_path & (dc <= 1e-4)
    both_ok = (out['ok'][sel] == 1) & (ref['ok'] == 1)
    w = np.asarray(cfg.weights, dtype=float)
    fd, fr = out['costs'][sel] @ w, ref['costs'] @ w
    rel = np.abs(fd - fr) / np.maximum(np.abs(fr), 1e-300)
    diff = ~good & both_ok
    print(f'M={M}: {int(good.sum())}/{n} problems identical to the checker (same attempt, status, iterations; coefficients <= 1e-4 m; '
          f'worst among them {dc[good].max():.2e} m); ok-flag agreement {np.mean(out["ok"][sel] == ref["ok"]):.5f}; of the {int((~good).sum())} others '
          f'{int(diff.sum())} are accepted by both sides with final cost ratio device/checker median {np.median(fd[diff] / fr[diff]) if diff.any() else 1:.4f} '
          f'(p10 {np.percentile(fd[diff] / fr[diff], 10) if diff.any() else 1:.3f}, p90 {np.percentile(fd[diff] / fr[diff], 90) if diff.any() else 1:.3f})')
    assert good.mean() >= floor
    # identical path, converged exit (the last evaluated point is the returned one, EP:233): the final costs agree as far
    # as decision vectors that differ in their last bits allow (cost and gradient AT IDENTICAL POINTS are compared to
    # 1e-6 in test_gpu_parity.py and to 1e-11 along whole runs in test_gpu_lockstep.py)
    conv = good & (ref['status'] <= 1)
    dx = np.max(np.abs(out['x'][sel] - ref['x']), axis=1)
    print(f'      converged among them: {int(conv.sum())}; worst |x_device - x_checker| {dx[conv].max():.1e}, worst relative cost difference '
          f'{rel[conv].max():.1e} (median {np.median(rel[conv]):.1e})')
    assert rel[conv].max() <= 1e-3 and np.median(rel[conv]) <= 1e-9
    return good


def derivs(c, T):
    """c (B,6,2), T (B,) -> list of the first five derivatives (B,2) of the quintic at T."""
    out = []
    for order in range(5):
        v = np.zeros((c.shape[0], 2))
        for k in range(order, 6):
            f = 1.0
            for q in range(order):
                f *= (k - q)
            v += f * c[:, k, :] * (T ** (k - order))[:, None]
        out.append(v)
    return out


def check_invariants(cfg, M, head, tail, out):
    ok = out['ok'] == 1
    c = out['coeffs'][ok].reshape(-1, M, 6, 2); ts = out['ts'][ok]
    h = lib.pad_state(head)[ok]; t = lib.pad_state(tail)[ok]
    assert ((ts > cfg.T_min) & (ts < cfg.T_max)).all()
    scale = 1.0 + np.abs(c).max()
    # head / tail states (EP:274-279, EP:318-332)
    assert np.abs(c[:, 0, 0] - h[:, 0]).max() < 1e-9 and np.abs(c[:, 0, 1] - h[:, 1]).max() < 1e-9
    assert np.abs(2 * c[:, 0, 2] - h[:, 2]).max() < 1e-9
    end = derivs(c[:, M - 1], ts[:, M - 1])
    for k in range(3):
        assert np.abs(end[k] - t[:, k]).max() < 1e-7 * scale
    # waypoints: position = q_i and continuity of vel, acc, jerk, snap (EP:284-316)
    x = out['x'][ok]
    for i in range(M - 1):
        e = derivs(c[:, i], ts[:, i]); s = derivs(c[:, i + 1], np.zeros(len(ts)))
        q = np.stack([x[:, i], x[:, (M - 1) + i]], axis=1)
        assert np.abs(e[0] - q).max() < 1e-7 * scale
        for k in range(5):
            assert np.abs(e[k] - s[k]).max() < 1e-6 * scale, (i, k)
    assert (out['costs'][ok][:, 3] * cfg.weights[3] <= cfg.collision_cost_tol).all()
    assert (out['nfev'] >= out['runs']).all() and (out['attempt'] < 5).all()


def test_config4_65536_problems_256_worlds():
    cfg = YamlConfig(); M = 3
    n_worlds, per = 256, 256
    h = pinned_handle(cfg, n_worlds, 8)          # 4 problems per warp: what a batch of this size runs by default
    heads, tails, worlds = [], [], []
    for wid in range(n_worlds):
        w = make_world(wid)
        worlds.append(w)
        h.set_map_occupancy(wid, w.H, w.W, w.res, w.ox, w.oy, w.occ)
        a, b = make_problems(w, per)
        heads.append(a); tails.append(b)
    head = np.concatenate(heads); tail = np.concatenate(tails); ids = np.repeat(np.arange(n_worlds, dtype=np.int32), per)
    B = len(head)
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(4))
    out = h.optimize(M, q0, ts0, head, tail, ids, rq, rts, 5)
    assert out['ok'].mean() > 0.85
    check_invariants(cfg, M, head, tail, out)
    # every one of the 65,536 problems against the CPU checker
    compare_with_checker(cfg, M, worlds, ids, head, tail, q0, ts0, rq, rts, out, np.arange(B), 0.985)
    # trajectory sampling for all 65,536 problems in one call (the batch index used to sit on gridDim.y <= 65,535)
    okm = out['ok'] == 1
    states, count = h.sample(M, out['coeffs'], np.where(okm[:, None], out['ts'], 1.0), 10.0)
    assert states.shape[0] == B and (count[okm] == np.ceil(out['ts'][okm].sum(1) * 10.0 - 1e-12).astype(int)).all()
    assert np.abs(states[okm, 0, 0] - head[okm][:, 0]).max() < 1e-9                  # first sample = start position
    last = B - 1 - int(np.argmax(okm[::-1]))
    assert np.abs(states[last, 0, 0] - head[last, 0]).max() < 1e-9 and count[last] > 10
    # determinism: warps pick tasks in a different order every launch; results must not depend on it
    again = h.optimize(M, q0, ts0, head, tail, ids, rq, rts, 5)
    for k in ('x', 'ts', 'coeffs', 'costs', 'status', 'ok', 'attempt', 'nit', 'runs', 'nfev'):
        assert np.array_equal(out[k], again[k]), k
    # independence of batch order and composition: a shuffled 4,096-problem subset gives the same per-problem results
    rng = np.random.default_rng(0)
    sub = rng.permutation(B)[:4096]
    part = h.optimize(M, q0[sub], ts0[sub], head[sub], tail[sub], ids[sub], rq[sub], rts, 5)
    for k in ('x', 'coeffs', 'costs', 'status', 'ok', 'attempt', 'nit', 'runs', 'nfev'):
        assert np.array_equal(part[k], out[k][sub]), k
    # speculative retries keep sequential semantics: where attempt 0 is accepted, a 1-attempt run returns the same thing
    # one problem per warp (what a small batch runs) on the same subset: same algorithm, different partial-sum order
    h32 = pinned_handle(cfg, n_worlds, 32)
    for wid in sorted(set(ids[sub].tolist())):
        w = worlds[wid]
        h32.set_map_occupancy(wid, w.H, w.W, w.res, w.ox, w.oy, w.occ)
    wide = h32.optimize(M, q0[sub], ts0[sub], head[sub], tail[sub], ids[sub], rq[sub], rts, 5)
    agree = ((wide['ok'] == part['ok']) & (wide['nit'] == part['nit']) & (wide['attempt'] == part['attempt'])
             & (np.max(np.abs(wide['coeffs'] - part['coeffs']).reshape(len(sub), -1), axis=1) <= 1e-4))
    print(f'8 vs 32 lanes per problem: {int(agree.sum())}/{len(sub)} identical outcomes')
    assert agree.mean() >= 0.985
    one = h.optimize(M, q0, ts0, head, tail, ids, max_attempts=1)
    first = out['attempt'] == 0
    assert first.mean() > 0.7
    for k in ('x', 'coeffs', 'costs', 'status', 'nit', 'nfev'):
        assert np.array_equal(one[k][first], out[k][first]), k
    assert np.array_equal(one['ok'][~first & (out['ok'] == 1)], np.zeros((~first & (out['ok'] == 1)).sum(), np.int32))


def test_config5_dense_map_10_pieces():
    cfg = YamlConfig(); M = 10; cfg.init_wpts_num = M - 1
    B = 16384
    w = make_world(0, dense=True)
    h = lib.Handle(cfg, 0, 1)
    h.set_map_occupancy(0, w.H, w.W, w.res, w.ox, w.oy, w.occ)
    head, tail = make_problems(w, B, M=M)
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(5))
    out = h.optimize(M, q0, ts0, head, tail, None, rq, rts, 5)
    assert out['ok'].mean() > 0.7
    check_invariants(cfg, M, head, tail, out)
    # every one of the 16,384 problems against the CPU checker (~8 ms per plan and core at this size)
    compare_with_checker(cfg, M, [w], None, head, tail, q0, ts0, rq, rts, out, np.arange(B), 0.97)
    again = h.optimize(M, q0, ts0, head, tail, None, rq, rts, 5)
    for k in ('x', 'coeffs', 'status', 'ok', 'attempt', 'nit', 'nfev'):
        assert np.array_equal(out[k], again[k]), k
    # cost/gradient linearity property of the map lookup: evaluating the returned x reproduces the stored costs on
    # converged exits (the last evaluated point is the returned one, EP:233)
    conv = (out['ok'] == 1) & (out['status'] <= 1)
    ev = h.eval(M, out['x'][conv], head[conv], tail[conv])
    assert np.allclose(ev['costs'], out['costs'][conv], rtol=1e-9, atol=1e-12)


def handle_with_env(cfg, n_maps, **env):
    """A handle created under development switches (read once, at neo_create)."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return lib.Handle(cfg, 0, n_maps)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.mark.parametrize('M,tile', [(2, 8), (3, 8), (3, 32), (4, 16), (6, 32), (10, 32)])
def test_grouped_starts_change_nothing(M, tile):
    """The one-CTA-per-SM variant (warps start their evaluations in groups on named barriers) against the plain one:
    identical outputs problem by problem, for batch sizes that leave most warps without work (1, 7), end inside a group
    (37 problems) or fill several SMs (1500) -- the drain path of the ticket/barrier scheme is what these exercise."""
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    w = make_world(3, dense=(M == 10))
    for B in (1, 7, 37, 1500):
        head, tail = make_problems(w, B, M=M, first=50 + B)
        q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
        rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(B))
        outs = []
        for grouped, group_warps in ((0, 0), (1, 9), (1, 12), (1, 2)):
            env = dict(NEO_TILE=tile, NEO_GROUPED=grouped)
            if group_warps:
                env['NEO_GROUP_WARPS'] = group_warps
            h = handle_with_env(cfg, 1, **env)
            h.set_map_occupancy(0, w.H, w.W, w.res, w.ox, w.oy, w.occ)
            outs.append(h.optimize(M, q0, ts0, lib.pad_state(head), lib.pad_state(tail), None, rq, rts, 5))
            h.close()
        for o in outs[1:]:
            for k in ('x', 'ts', 'coeffs', 'costs', 'status', 'ok', 'attempt', 'nit', 'runs', 'nfev'):
                assert np.array_equal(o[k], outs[0][k]), (B, k)

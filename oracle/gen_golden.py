"""gen_golden.py -- generates tests/golden/*.npz by importing and running the UNMODIFIED reference
from /root/reference (read-only) in the build container. Test infrastructure only.

    python oracle/gen_golden.py            # rewrites tests/golden/

While generating it also asserts that oracle/minco_ref.py (the Python restatement that travels to
the GPU box) is bit-identical to the reference on every vector written. The fixtures record the
numpy/scipy versions because scipy's L-BFGS-B and numpy.linalg.solve are un-pinned third-party
dependencies of the reference (README.md:43).
"""
import contextlib
import hashlib
import io
import os
import sys
import warnings

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/src/planner/scripts'
sys.path.insert(0, os.path.join(REF, 'traj_planner'))
sys.path.insert(0, os.path.join(REF, 'map_server'))
sys.path.insert(0, ROOT)

from expert_planner import MinJerkPlanner, DefaultConfig  # noqa: E402  (reference)
from esdf import ESDF  # noqa: E402  (reference)

from neo_planner_b200.worlds import make_world, make_problems, YamlConfig  # noqa: E402
from oracle import minco_ref  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
VERS = np.array([np.__version__, scipy.__version__])
warnings.filterwarnings('ignore')


class _EqSafe(np.ndarray):
    """numpy >= 2 raises on ``ndarray == []`` (TU:93 `if self.coeffs == []`); older numpy returned False.
    Viewing coeffs through this subclass restores the old answer so the reference's getters run unmodified."""
    def __eq__(self, other):
        if isinstance(other, list) and len(other) == 0:
            return False
        return np.ndarray.__eq__(self, other)
    __hash__ = None


class SamplingPlanner(MinJerkPlanner):
    def get_coeffs(self, int_wpts, ts):
        super().get_coeffs(int_wpts, ts)
        self.coeffs = self.coeffs.view(_EqSafe)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def ref_map(world):
    e = ESDF()
    e.occupancy_map_cb(world.occupancy_msg())
    return e


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cfg_for(M, base=YamlConfig):
    c = base()
    c.init_wpts_num = M - 1
    return c


def gen_esdf():
    """a20/a21: map build + lookups."""
    rng = np.random.default_rng(7)
    H, W, res, ox, oy = 48, 64, 0.1, -1.3, 2.05
    occ = np.where(rng.random((H, W)) < 0.03, 100, 0).astype(np.int8)
    occ[rng.random((H, W)) < 0.02] = -1            # unknown cells are free (ESDF:23)
    from types import SimpleNamespace as NS
    msg = NS(data=occ.reshape(-1).tolist(),
             info=NS(resolution=res, width=W, height=H, origin=NS(position=NS(x=ox, y=oy, z=0.0))))
    e = ESDF()
    e.occupancy_map_cb(msg)
    g = minco_ref.GridMap(occ, H, W, res, ox, oy)
    assert np.array_equal(g.esdf, e.esdf_map) and np.array_equal(g.gx, e.esdf_grad_x) and np.array_equal(g.gy, e.esdf_grad_y)
    # query points: inside, on cell borders, slightly outside (the (-1,0) truncation quirk), far outside
    pts = np.concatenate([
        rng.uniform([ox - 0.3, oy - 0.3], [ox + W * res + 0.3, oy + H * res + 0.3], size=(400, 2)),
        np.stack([ox + res * rng.integers(-2, W + 2, 100), oy + res * rng.integers(-2, H + 2, 100)], 1),
        np.array([[ox - 0.05, oy - 0.05], [ox - 0.0999, oy + 1.0], [ox + W * res, oy], [ox + 1.0, oy + H * res - 1e-12]]),
    ])
    dis = np.array([e.get_edt_dis(p) for p in pts], dtype=np.float64)
    grd = np.array([e.get_edt_grad(p) for p in pts], dtype=np.float64)
    idx = []
    for p in pts:
        r = int((p[1] - oy) / res)
        c = int((p[0] - ox) / res)
        idx.append((r, c) if (0 <= r < H and 0 <= c < W) else (-1, -1))
    for p, d0, g0 in zip(pts, dis, grd):
        assert g.get_edt_dis(p) == d0 and list(g.get_edt_grad(p)) == list(g0)
    # the all-free map (scipy's implementation-defined answer) and world 0 / dense world hashes
    free = np.zeros((9, 11), dtype=np.int8)
    e2 = ESDF()
    e2.occupancy_map_cb(NS(data=free.reshape(-1).tolist(), info=NS(resolution=0.2, width=11, height=9,
                                                                     origin=NS(position=NS(x=0.0, y=0.0, z=0.0)))))
    w0 = make_world(0)
    e0 = ref_map(w0)
    np.savez_compressed(os.path.join(OUT, 'esdf_small.npz'), versions=VERS,
                        occ=occ, H=H, W=W, res=res, ox=ox, oy=oy,
                        esdf=e.esdf_map, gx=e.esdf_grad_x, gy=e.esdf_grad_y,
                        pts=pts, dis=dis, grad=grd, idx=np.array(idx, dtype=np.int32),
                        free_esdf=e2.esdf_map, free_gx=e2.esdf_grad_x, free_gy=e2.esdf_grad_y,
                        world0_sha=np.array([sha(e0.esdf_map), sha(e0.esdf_grad_x), sha(e0.esdf_grad_y)]),
                        world0_occ_sha=np.array([sha(w0.occ)]))
    print('esdf_small.npz')


def gen_eval(M, n_prob, name, base=YamlConfig, world_id=0):
    """a12-a19: cost[4], cost, grad at probe points (x0, perturbed x0, optimum of attempt 0)."""
    cfg = cfg_for(M, base)
    w = make_world(world_id)
    e = ref_map(w)
    g = minco_ref.GridMap(w.occ, w.H, w.W, w.res, w.ox, w.oy)
    head, tail = make_problems(w, n_prob, M=M, v_max=cfg.v_max, safe_dis=cfg.safe_dis)
    pl = MinJerkPlanner(cfg)
    mine = minco_ref.RefOptimizer(cfg)
    rng = np.random.default_rng(1234 + M)
    X, HD, TL, C4, F, GR, CO = [], [], [], [], [], [], []
    for i in range(n_prob):
        iw, ts = pl.generate_init_variables(head[i], tail[i])
        if base is not YamlConfig:
            ts = ts + 0.7        # library defaults have init_T == T_min (log domain error); shift inside
        pl.read_planning_conditions(e, head[i], tail[i], iw, ts)
        mine.set_problem(g, head[i], tail[i], iw, ts)
        x0 = np.concatenate((iw.reshape(-1), pl.map_T2tau(ts)))
        probes = [x0]
        for _ in range(3):
            dx = np.concatenate((rng.normal(0, 0.4, 2 * (M - 1)), rng.normal(0, 0.8, M)))
            probes.append(x0 + dx)
        try:
            with quiet():
                pl.plan_once()
            probes.append(np.concatenate((pl.int_wpts.reshape(-1), pl.tau)))
        except Exception:
            pass
        pl.read_planning_conditions(e, head[i], tail[i], iw, ts)
        for x in probes:
            f = pl.get_cost(x)
            c4 = pl.costs.copy()
            gr = pl.get_grad(x)
            assert mine.cost(x) == f and np.array_equal(mine.costs, c4) and np.array_equal(mine.grad(x), gr)
            X.append(x); HD.append(pl.head_state.copy()); TL.append(pl.tail_state.copy())
            C4.append(c4); F.append(f); GR.append(gr); CO.append(pl.coeffs.copy())
    np.savez_compressed(os.path.join(OUT, name), versions=VERS, M=M, world_id=world_id,
                        cfg=np.array([cfg.v_max, cfg.T_min, cfg.T_max, cfg.safe_dis, cfg.delta_t, *cfg.weights,
                                      cfg.collision_cost_tol, cfg.init_T], dtype=np.float64),
                        x=np.array(X), head=np.array(HD), tail=np.array(TL), costs=np.array(C4),
                        f=np.array(F), grad=np.array(GR), coeffs=np.array(CO))
    print(name, len(X), 'probe points; collision>0 at', int((np.array(C4)[:, 3] > 0).sum()),
          'feas>0 at', int((np.array(C4)[:, 2] > 0).sum()))


def gen_plans(M, n_prob, name, world_id=0):
    """a9/a10 + a23: plan_once from the expert guess, and full plan() with retries (np.random.seed(k))."""
    cfg = cfg_for(M)
    w = make_world(world_id)
    e = ref_map(w)
    g = minco_ref.GridMap(w.occ, w.H, w.W, w.res, w.ox, w.oy)
    head, tail = make_problems(w, n_prob, M=M)
    n = 2 * (M - 1) + M
    x0s = np.zeros((n_prob, n)); xs = np.zeros((n_prob, n)); nit = np.zeros(n_prob, np.int32)
    nfev = np.zeros(n_prob, np.int32); msg = []; c4 = np.zeros((n_prob, 4)); exc = []
    # full plan() outputs
    P_ok = np.zeros(n_prob, np.int32); P_x = np.zeros((n_prob, n)); P_ts = np.zeros((n_prob, M))
    P_coeffs = np.zeros((n_prob, 6 * M, 2)); P_iter = np.zeros(n_prob, np.int32); P_runs = np.zeros(n_prob, np.int32)
    P_costs = np.zeros((n_prob, 4)); P_cmd_sha = []; P_cmd0 = None
    for i in range(n_prob):
        pl = MinJerkPlanner(cfg)
        mine = minco_ref.RefOptimizer(cfg)
        iw, ts = pl.generate_init_variables(head[i], tail[i])
        pl.read_planning_conditions(e, head[i], tail[i], iw, ts)
        mine.set_problem(g, head[i], tail[i], iw.copy(), ts.copy())
        x0s[i] = np.concatenate((iw.reshape(-1), pl.map_T2tau(ts)))
        import scipy.optimize as sopt
        try:
            res = sopt.minimize(pl.get_cost, x0s[i], method='L-BFGS-B', jac=pl.get_grad, bounds=None, tol=1e-4,
                                options={'maxcor': 10, 'maxfun': 15000, 'maxiter': 15000, 'maxls': 20})
            xs[i] = res.x; nit[i] = res.nit; nfev[i] = res.nfev; msg.append(res.message); c4[i] = pl.costs
            exc.append('')
        except Exception as ex:
            msg.append('EXC'); exc.append(type(ex).__name__)
        # full plan with reproducible retry noise
        pl = SamplingPlanner(cfg)
        np.random.seed(i)
        try:
            with quiet():
                pl.plan(e, head[i], tail[i])
            P_ok[i] = 1
            P_x[i] = np.concatenate((pl.int_wpts.reshape(-1), pl.tau)); P_ts[i] = pl.ts
            P_costs[i] = pl.costs
            cmd = pl.get_full_state_cmd(60)
            P_coeffs[i] = np.asarray(pl.coeffs)
            P_cmd_sha.append(sha(cmd))
            if P_cmd0 is None:
                P_cmd0 = (i, cmd)
        except Exception:
            P_cmd_sha.append('')
        P_iter[i] = pl.iter_num; P_runs[i] = pl.opt_running_times
        # python restatement must reproduce the reference bit for bit (same RNG stream)
        mine2 = minco_ref.RefOptimizer(cfg)
        np.random.seed(i)
        try:
            mine2.plan(g, head[i], tail[i])
            ok2 = 1
        except Exception:
            ok2 = 0
        assert ok2 == P_ok[i] and mine2.iter_num == P_iter[i] and mine2.opt_running_times == P_runs[i], (i,)
        if ok2:
            assert np.array_equal(np.concatenate((mine2.int_wpts.reshape(-1), mine2.tau)), P_x[i])
            assert sha(mine2.full_state_cmd(60)) == P_cmd_sha[-1]
    np.savez_compressed(os.path.join(OUT, name), versions=VERS, M=M, world_id=world_id,
                        head=head, tail=tail, x0=x0s, x=xs, nit=nit, nfev=nfev, msg=np.array(msg), exc=np.array(exc),
                        costs=c4, plan_ok=P_ok, plan_x=P_x, plan_ts=P_ts, plan_coeffs=P_coeffs, plan_iter=P_iter,
                        plan_runs=P_runs, plan_costs=P_costs, cmd_index=P_cmd0[0], cmd=P_cmd0[1])
    print(name, 'messages:', {m: msg.count(m) for m in set(msg)}, 'plan ok', int(P_ok.sum()), '/', n_prob,
          'mean runs', P_runs.mean())


def gen_batch_plan(n_prob, name):
    """a6/a7: three-candidate batch_plan."""
    cfg = cfg_for(3)
    w = make_world(1)
    e = ref_map(w)
    g = minco_ref.GridMap(w.occ, w.H, w.W, w.res, w.ox, w.oy)
    head, tail = make_problems(w, n_prob, M=3)
    cands = np.zeros((n_prob, 3, 2, 2)); X = np.zeros((n_prob, 4)); TS = np.zeros((n_prob, 3)); FC = np.zeros(n_prob)
    OK = np.zeros(n_prob, np.int32)
    for i in range(n_prob):
        pl = MinJerkPlanner(cfg)
        cands[i], ts0 = pl.batch_generate_init_variables(head[i], tail[i])
        np.random.seed(100 + i)
        try:
            with quiet():
                pl.batch_plan(e, head[i], tail[i])
            OK[i] = 1; X[i] = pl.int_wpts.reshape(-1); TS[i] = pl.ts; FC[i] = pl.final_cost
        except Exception:
            pass
        mine = minco_ref.RefOptimizer(cfg)
        np.random.seed(100 + i)
        try:
            mine.batch_plan(g, head[i], tail[i])
            assert OK[i] and np.array_equal(mine.int_wpts.reshape(-1), X[i]) and np.array_equal(mine.ts, TS[i])
        except AssertionError:
            raise
        except Exception:
            assert not OK[i]
    np.savez_compressed(os.path.join(OUT, name), versions=VERS, head=head, tail=tail, cands=cands, ts0=ts0,
                        ok=OK, int_wpts=X, ts=TS, final_cost=FC)
    print(name, 'ok', int(OK.sum()), '/', n_prob)


def gen_errors(name):
    """Q11 / a11: error behaviour. Library defaults have init_T == T_min -> ZeroDivisionError on every attempt."""
    w = make_world(0)
    e = ref_map(w)
    head, tail = make_problems(w, 2, M=3)
    pl = MinJerkPlanner(DefaultConfig())
    np.random.seed(0)
    try:
        with quiet():
            pl.plan(e, head[0], tail[0])
        outcome = 'ok'
    except Exception as ex:
        outcome = str(ex)
    # ts outside (T_min, T_max) as an NN guess could produce (EP:209)
    pl2 = MinJerkPlanner(YamlConfig())
    iw, ts = pl2.generate_init_variables(head[1], tail[1])
    bad = ts.copy(); bad[1] = 5.5
    np.random.seed(3)
    with quiet():
        pl2.warm_start_plan(e, head[1], tail[1], iw, bad)
    np.savez_compressed(os.path.join(OUT, name), versions=VERS, head=head, tail=tail,
                        default_outcome=np.array([outcome]), default_runs=pl.opt_running_times,
                        bad_ts=bad, bad_x=np.concatenate((pl2.int_wpts.reshape(-1), pl2.tau)), bad_runs=pl2.opt_running_times,
                        bad_iter=pl2.iter_num)
    print(name, outcome, pl.opt_running_times, pl2.opt_running_times)


def gen_geo(name):
    """§8f rank 4: AstarPlanner.plan (astar_planner.py:22-103) + GeoPlanner.prune_path_nodes / geo_traj_plan
    (geo_planner.py:19-101) from the unmodified reference. astar_planner imports matplotlib (absent here) only for its
    unused visualize_path helper, so an empty stand-in module is registered before the import."""
    import types
    from types import SimpleNamespace as NS
    for mod in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    from astar_planner import AstarPlanner  # noqa: E402  (reference)
    from geo_planner import GeoPlanner  # noqa: E402  (reference)
    from oracle import astar_ref
    cfg = cfg_for(3)
    sets = []
    for world_id, dense, Mlen, n in ((0, False, 3, 24), (0, False, 10, 16), (2, False, 10, 8), (1, True, 10, 4)):
        w = make_world(world_id, dense=dense)
        e = ref_map(w)
        g = minco_ref.GridMap(w.occ, w.H, w.W, w.res, w.ox, w.oy)
        head, tail = make_problems(w, n, M=Mlen)
        sets.append((world_id, dense, w, e, g, head, tail))
    wid = []; dns = []; heads = []; tails = []; paths = []; plen = []; pruned = []
    P_ok = []; P_x = []; P_ts = []; P_iter = []; P_runs = []; n_keys = []
    for world_id, dense, w, e, g, head, tail in sets:
        for i in range(head.shape[0]):
            ap = AstarPlanner()
            path = ap.plan(e, head[i, 0], tail[i, 0])
            gp = GeoPlanner(cfg)
            four = gp.prune_path_nodes(e, path)
            mine, found, _ = astar_ref.astar(g, head[i, 0], tail[i, 0])
            assert found and mine == path, (world_id, i)
            if len(path) < 80:
                assert astar_ref.astar_plain(g, head[i, 0], tail[i, 0]) == path
            four2, pick, keys = astar_ref.prune(g, path)
            assert four2 == four, (world_id, i)
            n_keys.append(len(keys))
            wid.append(world_id); dns.append(int(dense)); heads.append(head[i]); tails.append(tail[i])
            paths.append(np.array(path)); plen.append(len(path)); pruned.append(np.array(four))
            # the full geometric warm start (GEO:19-39) with reproducible retry noise
            if not dense:
                st = NS(global_pos=np.array([head[i, 0, 0], head[i, 0, 1], 2.0]),
                        global_vel=np.array([head[i, 1, 0], head[i, 1, 1], 0.0]))
                gp = GeoPlanner(cfg)
                np.random.seed(100 + i)
                try:
                    with quiet():
                        gp.geo_traj_plan(e, st, tail[i])
                    P_ok.append(1); P_x.append(np.concatenate((gp.int_wpts.reshape(-1), gp.tau))); P_ts.append(gp.ts)
                except Exception:
                    P_ok.append(0); P_x.append(np.zeros(7)); P_ts.append(np.zeros(3))
                P_iter.append(gp.iter_num); P_runs.append(gp.opt_running_times)
            else:
                P_ok.append(-1); P_x.append(np.zeros(7)); P_ts.append(np.zeros(3)); P_iter.append(0); P_runs.append(0)
    # an unreachable target (inside a pillar): the reference exhausts the grid and returns [target cell] (AP:58-60).
    # Only feasible to run on a tiny map, the reference's open-set scan is quadratic.
    occ = np.zeros((12, 16), np.int8); occ[5:8, 9:12] = 100
    tiny = NS(info=NS(resolution=1.0, width=16, height=12, origin=NS(position=NS(x=0.0, y=0.0))), data=occ.reshape(-1))
    et = ESDF(); et.occupancy_map_cb(tiny)
    ap = AstarPlanner()
    with quiet():
        lost = ap.plan(et, [2.5, 2.5], [10.5, 6.5])
    gt = minco_ref.GridMap(occ, 12, 16, 1.0, 0.0, 0.0)
    mine, found, nclosed = astar_ref.astar(gt, [2.5, 2.5], [10.5, 6.5])
    assert not found and mine == lost
    lost_four = GeoPlanner(cfg).prune_path_nodes(et, lost)
    assert astar_ref.prune(gt, lost)[0] == lost_four
    np.savez_compressed(os.path.join(OUT, name), versions=VERS, world_id=np.array(wid), dense=np.array(dns),
                        head=np.array(heads), tail=np.array(tails), path=np.concatenate(paths), path_len=np.array(plen),
                        pruned=np.array(pruned), n_keys=np.array(n_keys), plan_ok=np.array(P_ok), plan_x=np.array(P_x),
                        plan_ts=np.array(P_ts), plan_iter=np.array(P_iter), plan_runs=np.array(P_runs),
                        tiny_occ=occ, lost_path=np.array(lost), lost_pruned=np.array(lost_four), lost_closed=nclosed)
    print(name, 'paths', len(plen), 'len', min(plen), max(plen), 'key-node counts', sorted(set(n_keys)),
          'geo plans ok', sum(1 for v in P_ok if v == 1), '/', sum(1 for v in P_ok if v >= 0))


def stub_modules(*names):
    """Registers empty stand-in modules (with a module spec, torch's import scans ask for it) for packages the reference
    imports but never uses on the paths exercised here."""
    import importlib.machinery
    import types
    for nm in names:
        if nm not in sys.modules:
            m = types.ModuleType(nm)
            m.__spec__ = importlib.machinery.ModuleSpec(nm, None)
            sys.modules[nm] = m
    return [sys.modules[nm] for nm in names]


class Quaternion:
    """Stand-in for pyquaternion.Quaternion (not installed here; the reference's record_planner / nn_planner use
    `.rotate`, `.inverse`, `.rotation_matrix` of it). Same algebra as pyquaternion 0.9: products through the 4x4 left
    matrix, rotation as q v q*, rotation matrix as the lower-right 3x3 of Q(q) Qbar(q)^T, each after normalisation."""

    def __init__(self, w=1.0, x=0.0, y=0.0, z=0.0, array=None):
        self.q = np.array([w, x, y, z], dtype=float) if array is None else np.array(array, dtype=float)

    def _normalised(self):
        n = np.sqrt(np.dot(self.q, self.q))
        return Quaternion(array=self.q / n) if abs(1.0 - n) > 1e-14 else self

    def _q_matrix(self):
        w, x, y, z = self.q
        return np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])

    def _q_bar_matrix(self):
        w, x, y, z = self.q
        return np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]])

    def __mul__(self, other):
        return Quaternion(array=np.dot(self._q_matrix(), other.q))

    @property
    def conjugate(self):
        return Quaternion(array=self.q * np.array([1.0, -1.0, -1.0, -1.0]))

    @property
    def inverse(self):
        return Quaternion(array=self.conjugate.q / np.dot(self.q, self.q))

    def rotate(self, v):
        u = self._normalised()
        return (u * Quaternion(array=np.concatenate(([0.0], np.asarray(v, dtype=float)))) * u.conjugate).q[1:]

    @property
    def rotation_matrix(self):
        u = self._normalised()
        return np.dot(u._q_matrix(), u._q_bar_matrix().conj().transpose())[1:][:, 1:]


def gen_nn_io(name):
    """§8f rank 3 / a24: form_nn_input, form_nn_output (record_planner.py:13-72) and NNPlanner.get_wpts_world
    (nn_planner.py:123-134) of the unmodified reference on seeded drone states. pyquaternion and onnxruntime are not
    installed: the Quaternion stand-in above and empty onnxruntime / torchinfo / onnx / matplotlib modules are registered
    before the imports (nn_planner is only imported for the get_wpts_world method, no session is created)."""
    import types
    from types import SimpleNamespace as NS
    stub_modules('onnxruntime', 'torchinfo', 'onnx', 'matplotlib', 'matplotlib.pyplot', 'pyquaternion')
    sys.modules['torchinfo'].summary = lambda *a, **k: None
    sys.modules['pyquaternion'].Quaternion = Quaternion
    sys.path.insert(0, REF)                                   # nn_planner imports nn_trainer.nn_trainer as a package
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    from record_planner import form_nn_input, form_nn_output  # noqa: E402  (reference)
    from nn_planner import NNPlanner  # noqa: E402  (reference)
    rng = np.random.default_rng(31)
    B = 12
    des_z = 2.0
    quat = rng.normal(size=(B, 4)); quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    quat[0] = [1, 0, 0, 0]; quat[1] = [np.cos(0.3), 0, 0, np.sin(0.3)]
    local_vel = rng.normal(0, 1, (B, 3)); gpos = rng.uniform(-5, 25, (B, 3)); gvel = rng.normal(0, 1, (B, 3))
    ipos = gpos + rng.normal(0, 0.5, (B, 3)); ivel = gvel + rng.normal(0, 0.2, (B, 3))
    target = np.stack([ipos[:, :2] + rng.uniform(3, 6, (B, 2)), rng.normal(0, 1, (B, 2))], axis=1)
    depth = rng.uniform(0.2, 12.0, (B, 48, 64)).astype(np.float32)
    int_wpts = rng.uniform(-5, 25, (B, 2, 2))
    net_local = rng.normal(0, 2, (B, 3, 2))
    d_norm = []; motion = []; out_local = []; world = []
    for k in range(B):
        q = Quaternion(*quat[k])
        ds = NS(local_vel=local_vel[k], attitude=q, global_pos=gpos[k], global_vel=gvel[k])
        st = NS(global_pos=ipos[k], global_vel=ivel[k])
        dn, mi = form_nn_input(depth[k], ds, des_z, st, target[k])
        d_norm.append(dn); motion.append(mi)
        out_local.append(form_nn_output(ds, des_z, int_wpts[k]))
        fake = NS(drone_state=ds, nn_output_D=3, M=3)
        world.append(NNPlanner.get_wpts_world(fake, net_local[k]))
    np.savez_compressed(os.path.join(OUT, name), versions=VERS, des_pos_z=des_z, quat=quat, local_vel=local_vel, global_pos=gpos,
                        global_vel=gvel, init_pos=ipos, init_vel=ivel, target=target, depth=depth, int_wpts=int_wpts,
                        net_local=net_local, depth_norm=np.array(d_norm), motion=np.array(motion),
                        int_wpts_local=np.array(out_local), wpts_world=np.array(world))
    print(name, 'samples', B)


def gen_nets(name):
    """a24 / §8f rank 4: the reference's two PlannerNet classes (nn_trainer.py:109-155 MLP heads, nn_trainer_conv.py:108-160
    Conv1d heads) instantiated unmodified -- torchvision's resnet18 is told not to download weights -- filled with
    oracle/net_weights.fill_deterministic and run in fp32 on oracle/net_weights.sample_input."""
    import types
    import torch
    import torchvision.models as tvm
    from oracle import net_weights
    stub_modules('onnxruntime', 'torchinfo', 'onnx', 'matplotlib', 'matplotlib.pyplot')
    sys.modules['torchinfo'].summary = lambda *a, **k: None
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    orig = tvm.resnet18
    tvm.resnet18 = lambda weights=None, **kw: orig(weights=None, **kw)       # no network here: skip the ImageNet download
    import importlib.util

    def load(fname):                                     # the unmodified reference file, loaded by path
        spec = importlib.util.spec_from_file_location('ref_' + fname[:-3], os.path.join(REF, 'nn_trainer', fname))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    try:
        ref_mlp = load('nn_trainer.py')
        ref_conv = load('nn_trainer_conv.py')
        x = torch.from_numpy(net_weights.sample_input())
        out = {}
        for tag, mod in (('mlp', ref_mlp), ('conv', ref_conv)):
            net = mod.PlannerNet().eval()
            net_weights.fill_deterministic(net)
            with torch.no_grad():
                y = net(x)
            out[tag + '_out'] = y.reshape(x.shape[0], -1).numpy()
            out[tag + '_sig'] = np.array(net_weights.signature(net))
    finally:
        tvm.resnet18 = orig
    np.savez_compressed(os.path.join(OUT, name), versions=VERS, torch_version=np.array(torch.__version__), **out)
    print(name, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:      # python oracle/gen_golden.py nn_io nets  -> only these fixtures
        for which in sys.argv[1:]:
            {'nn_io': lambda: gen_nn_io('nn_io.npz'), 'nets': lambda: gen_nets('nets.npz')}[which]()
        sys.exit(0)
    gen_esdf()
    gen_eval(3, 24, 'eval_M3.npz')
    gen_eval(10, 8, 'eval_M10.npz')
    gen_eval(3, 8, 'eval_M3_libdefaults.npz', base=DefaultConfig)
    gen_plans(3, 96, 'plans_M3.npz')
    gen_plans(10, 24, 'plans_M10.npz')
    gen_batch_plan(16, 'batch_plan_M3.npz')
    gen_errors('errors.npz')
    gen_geo('geo_M3.npz')
    gen_nn_io('nn_io.npz')
    gen_nets('nets.npz')

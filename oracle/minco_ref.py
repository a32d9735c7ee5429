"""minco_ref.py -- TEST INFRASTRUCTURE ONLY (not shipped, not on the product path).

Python/NumPy/SciPy restatement of the reference NEO-Planner hot path, used
  (1) as the checker in tests/ (it calls the *real* scipy L-BFGS-B, like the reference does), and
  (2) as the `--impl reference` / `cpu_baseline` arm of bench.py on the GPU box, where
      /root/reference does not exist.

It follows, function by function (paths relative to /root/reference):
  EP   = src/planner/scripts/traj_planner/expert_planner.py
  TU   = src/planner/scripts/traj_planner/traj_utils.py
  ESDF = src/planner/scripts/map_server/esdf.py
and keeps the reference's NumPy call shapes (np.dot on the same strided views, ``**`` powers,
Python ``sum``) so that results are bit-identical to the reference on the same machine.
Parity pin: tests/test_oracle_golden.py compares this module with tests/golden/ fixtures that
oracle/gen_golden.py produced by importing and running the unmodified reference in the build
container (numpy 2.3.5, scipy 1.18.1); gen_golden.py itself asserts bit-equality while generating.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
import scipy.optimize as sopt
from scipy import ndimage

S_ORDER = 3  # minimum-jerk: EP:36


class GridMap:
    """ESDF:11-82 restated. Build from raw OccupancyGrid values (100 = occupied)."""

    def __init__(self, occ, H, W, res, ox, oy):
        raw = np.asarray(occ).reshape(-1)
        binary = tuple(1 if v == 100 else 0 for v in raw)          # ESDF:23 (unknown -> free)
        self.H, self.W, self.res, self.ox, self.oy = int(H), int(W), res, ox, oy
        self.occ2d = np.array(binary).reshape(self.H, self.W)      # ESDF:26
        self.esdf = ndimage.distance_transform_edt(1 - self.occ2d) * res   # ESDF:29
        self.gy, self.gx = np.gradient(self.esdf)                  # ESDF:33

    def cell(self, pos):
        r = int((pos[1] - self.oy) / self.res)                     # ESDF:61
        c = int((pos[0] - self.ox) / self.res)                     # ESDF:62
        if r < 0 or r >= self.H or c < 0 or c >= self.W:
            return None
        return r, c

    def get_edt_dis(self, pos):                                    # ESDF:53-67
        rc = self.cell(pos)
        return 10000 if rc is None else self.esdf[rc]

    def get_edt_grad(self, pos):                                   # ESDF:69-82
        rc = self.cell(pos)
        return [0, 0] if rc is None else [self.gx[rc], self.gy[rc]]

    def has_collision(self, pos):                                  # ESDF:50 (module constant 0.5)
        return self.get_edt_dis(pos) < 0.5


def monomial_table(T_max, dt):
    """EP:250-259: rows [beta0..beta3](t_j), t_j = arange(0, T_max, dt)[j]."""
    tj = np.arange(0, T_max, dt)
    tab = np.zeros((len(tj), 4, 6))
    for j, t in enumerate(tj):
        tab[j, 0] = [1, t, t**2, t**3, t**4, t**5]
        tab[j, 1] = [0, 1, 2*t, 3*t**2, 4*t**3, 5*t**4]
        tab[j, 2] = [0, 0, 2, 6*t, 12*t**2, 20*t**3]
        tab[j, 3] = [0, 0, 0, 6, 24*t, 60*t**2]
    return tab


def minco_system(ts, head3, tail3, q):
    """EP:261-336 / TU:8-83. q: (D, M-1); head3/tail3: (3, D). Returns A (6M,6M), b (6M,D)."""
    M = ts.shape[0]
    D = head3.shape[1]
    n = 6 * M
    p1, p2, p3, p4, p5 = ts, ts**2, ts**3, ts**4, ts**5
    A = np.zeros((n, n))
    b = np.zeros((n, D))
    b[:3] = head3
    b[n-3:] = tail3
    A[0, 0] = 1.0
    A[1, 1] = 1.0
    A[2, 2] = 2.0
    wp = q.T
    for i in range(M - 1):
        r, c = 6*i + 3, 6*i
        row_p = [1.0, p1[i], p2[i], p3[i], p4[i], p5[i]]
        A[r, c:c+6] = row_p                                   # p_i(T_i) = q_i
        A[r+1, c:c+6] = row_p                                 # position continuity
        A[r+1, c+6] = -1.0
        A[r+2, c+1:c+6] = [1.0, 2*p1[i], 3*p2[i], 4*p3[i], 5*p4[i]]
        A[r+2, c+7] = -1.0
        A[r+3, c+2:c+6] = [2.0, 6*p1[i], 12*p2[i], 20*p3[i]]
        A[r+3, c+8] = -2.0
        A[r+4, c+3:c+6] = [6.0, 24.0*p1[i], 60.0*p2[i]]
        A[r+4, c+9] = -6.0
        A[r+5, c+4:c+6] = [24.0, 120.0*p1[i]]
        A[r+5, c+10] = -24.0
        b[r] = wp[i]
    A[n-3, n-6:] = [1.0, p1[-1], p2[-1], p3[-1], p4[-1], p5[-1]]
    A[n-2, n-5:] = [1.0, 2*p1[-1], 3*p2[-1], 4*p3[-1], 5*p4[-1]]
    A[n-1, n-4:] = [2.0, 6*p1[-1], 12*p2[-1], 20*p3[-1]]
    return A, b


def jerk_gram(T):
    """EP:352-358: integral of jerk^2 over [0,T] as a 6x6 quadratic form."""
    G = np.zeros((6, 6))
    G[3, 3:] = [36*T, 72*T**2, 120*T**3]
    G[4, 3:] = [72*T**2, 192*T**3, 360*T**4]
    G[5, 3:] = [120*T**3, 360*T**4, 720*T**5]
    return G


class Params:
    """The a1 parameter struct (EP:12-25 / planner_config.yaml:2-13)."""
    FIELDS = ('v_max', 'T_min', 'T_max', 'safe_dis', 'delta_t', 'weights', 'init_wpts_mode',
              'init_seg_len', 'init_wpts_num', 'init_T', 'collision_cost_tol', 'opt_tol')

    def __init__(self, cfg):
        for k in self.FIELDS:
            setattr(self, k, getattr(cfg, k))
        self.weights = np.array(cfg.weights)
        self.init_wpts_num = int(cfg.init_wpts_num)


class RefOptimizer:
    """One planning problem at a time, like the reference object (stateful; EP:28-60)."""

    def __init__(self, cfg):
        self.p = Params(cfg)
        self.table = monomial_table(self.p.T_max, self.p.delta_t)
        self.iter_num = 0
        self.opt_running_times = 0
        self.nfev = 0

    # ---- planning conditions (EP:170-184) -------------------------------------------------
    def set_problem(self, grid, head_state, tail_state, int_wpts, ts):
        self.map = grid
        self.D = head_state.shape[1]
        self.M = ts.shape[0]
        self.head = np.zeros((S_ORDER, self.D))
        self.tail = np.zeros((S_ORDER, self.D))
        for k in range(min(S_ORDER, head_state.shape[0])):
            self.head[k] = head_state[k]
        for k in range(min(S_ORDER, tail_state.shape[0])):
            self.tail[k] = tail_state[k]
        self.int_wpts = int_wpts
        self.ts = ts

    # ---- time reparametrisation (EP:468-492) ----------------------------------------------
    def T2tau(self, ts):
        p = self.p
        out = np.zeros(self.M)
        for i in range(self.M):
            out[i] = -math.log((p.T_max - p.T_min) / (ts[i] - p.T_min) - 1)
        return out

    def tau2T(self, tau):
        p = self.p
        out = np.zeros(self.M)
        for i in range(self.M):
            out[i] = (p.T_max - p.T_min) / (1 + math.exp(-tau[i])) + p.T_min
        return out

    # ---- initial guesses (EP:82-140) ------------------------------------------------------
    def straight_line_guess(self, head_state, tail_state, seed=0):
        p = self.p
        a, z = head_state[0], tail_state[0]
        if p.init_wpts_mode == 'adaptive':
            k = max(math.ceil(np.linalg.norm(z - a) / p.init_seg_len - 1), 1)
        else:
            k = p.init_wpts_num
        hop = (z - a) / (k + 1)
        w = np.linspace(a + hop, z, k, endpoint=False)
        if seed != 0:
            w += np.random.normal(0, 0.5, w.shape)      # global RNG, as EP:94
        ts = p.init_T * np.ones((k + 1,))
        ts[0] *= 1.5
        ts[-1] *= 1.5
        return w.T, ts

    def lateral_guesses(self, head_state, tail_state, count=3, shift=0.6):
        p = self.p
        a, z = head_state[0], tail_state[0]
        u = (z - a) / np.linalg.norm(z - a)
        side = np.array([[u[1], -u[0]], [-u[1], u[0]]])
        k = p.init_wpts_num
        cands = np.zeros((count, k, head_state.shape[1]))
        hop = (z - a) / (k + 1)
        cands[0] = np.linspace(a + hop, z, k, endpoint=False)
        flag = 0
        for j in range(1, count):
            cands[j] = cands[0] + shift * side[flag]
            flag = 1 - flag
        ts = p.init_T * np.ones((k + 1,))
        ts[0] *= 1.5
        ts[-1] *= 1.5
        return np.transpose(cands, (0, 2, 1)), ts

    # ---- cost / gradient (EP:345-466, EP:494-585) -----------------------------------------
    def _unpack(self, x):
        nq = self.D * (self.M - 1)
        self.int_wpts = np.reshape(x[:nq], (self.D, self.M - 1))
        self.tau = x[nq:]
        self.ts = self.tau2T(self.tau)
        self.A, rhs = minco_system(self.ts, self.head, self.tail, self.int_wpts)
        self.coeffs = np.linalg.solve(self.A, rhs)

    def cost(self, x):
        self._unpack(x)
        p, M = self.p, self.M
        self.costs = np.zeros(len(p.weights))
        for i in range(M):
            c = self.coeffs[6*i:6*(i+1), :]
            self.costs[0] += np.trace(c.T @ jerk_gram(self.ts[i]) @ c)
        self.costs[1] += np.sum(self.ts)
        for i in range(M):
            c = self.coeffs[6*i:6*(i+1), :]
            ns = int(self.ts[i] / p.delta_t)
            for j in range(ns):
                beta = self.table[j]
                pos = np.dot(c.T, beta[0])
                vel = np.dot(c.T, beta[1])
                omg = 0.5 if j in [0, ns - 1] else 1
                over_v = sum(vel**2) - p.v_max**2
                if over_v > 0.0:
                    self.costs[2] += omg * p.delta_t * over_v**3
                over_d = p.safe_dis - self.map.get_edt_dis(pos[:2])
                if over_d > 0.0:
                    self.costs[3] += omg * p.delta_t * over_d**3
        return np.dot(self.costs, p.weights)

    def grad(self, x):
        self._unpack(x)
        p, M, D = self.p, self.M, self.D
        gC = np.zeros((6 * M, D))
        gT = np.zeros(M)
        for i in range(M):
            c = self.coeffs[6*i:6*(i+1), :]
            T = self.ts[i]
            j3 = np.array([[0, 0, 0, 6, 24*T, 60*T**2]]).T
            gC[6*i:6*(i+1), :] += p.weights[0] * 2 * jerk_gram(T) @ c
            for d in range(D):
                gT[i] += p.weights[0] * (np.dot(c[:, d], j3).item())**2
        gT += p.weights[1] * np.ones(M)
        for i in range(M):
            c = self.coeffs[6*i:6*(i+1), :]
            ns = int(self.ts[i] / p.delta_t)
            for j in range(ns):
                beta = self.table[j]
                pos = np.dot(c.T, beta[0])
                vel = np.dot(c.T, beta[1])
                omg = 0.5 if j in [0, ns - 1] else 1
                over_v = sum(vel**2) - p.v_max**2
                if over_v > 0.0:
                    dv_dc = 2 * np.dot(np.array([beta[1]]).T, np.array([vel]))
                    dv_dt = 2 * (np.array([beta[2]]) @ c @ np.array([vel]).T).item()
                    K = 3 * p.delta_t * omg * over_v**2
                    gC[6*i:6*(i+1), :] += p.weights[2] * K * dv_dc
                    gT[i] += p.weights[2] * (omg*over_v**3/ns + K * dv_dt * j/ns)
                over_d = p.safe_dis - self.map.get_edt_dis(pos[:2])
                if over_d > 0.0:
                    eg = self.map.get_edt_grad(pos[:2])
                    K = 3 * p.delta_t * omg * over_d**2
                    dpsi_dc = -np.array([beta[0]]).T @ np.array([eg])
                    dpsi_dt = (-np.array([eg]) @ np.array([vel]).T).item()
                    gC[6*i:6*(i+1), :] += p.weights[3] * K * dpsi_dc
                    gT[i] += p.weights[3] * (omg*over_d**3/ns + K * dpsi_dt * j/ns)
        # adjoint (EP:494-537)
        G = np.linalg.solve(self.A.T, gC)
        gq = np.zeros((D, M - 1))
        gT2 = np.zeros(M)
        for i in range(M - 1):
            gq[:, i] = G[6*i + 3, :].T
        for i in range(M - 1):
            T = self.ts[i]
            dE = np.array([[0, 1, 2*T, 3*T**2, 4*T**3, 5*T**4],
                           [0, 1, 2*T, 3*T**2, 4*T**3, 5*T**4],
                           [0, 0, 2, 6*T, 12*T**2, 20*T**3],
                           [0, 0, 0, 6, 24*T, 60*T**2],
                           [0, 0, 0, 0, 24, 120*T],
                           [0, 0, 0, 0, 0, 120]])
            gT2[i] = gT[i] - np.trace(G[6*i+3:6*i+9, :].T @ dE @ self.coeffs[6*i:6*i+6, :])
        # last piece: the reference re-uses the loop variable T = ts[M-2] here (EP:527-533)
        dE = np.array([[0, 1, 2*T, 3*T**2, 4*T**3, 5*T**4],
                       [0, 0, 2, 6*T, 12*T**2, 20*T**3],
                       [0, 0, 0, 6, 24*T, 60*T**2]])
        gT2[-1] = gT[-1] - np.trace(G[-3:, :].T @ dE @ self.coeffs[-6:, :])
        gtau = np.zeros(M)
        for i in range(M):
            gtau[i] = gT2[i] * (p.T_max - p.T_min) * math.exp(-self.tau[i]) / (1 + math.exp(-self.tau[i]))**2
        return np.concatenate((np.reshape(gq, (D * (M - 1),)), gtau), axis=0)

    # ---- optimisation (EP:186-237) --------------------------------------------------------
    def plan_once(self):
        self.tau = self.T2tau(self.ts)
        nq = self.D * (self.M - 1)
        x0 = np.concatenate((np.reshape(self.int_wpts, (nq,)), self.tau), axis=0)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = sopt.minimize(self.cost, x0, method='L-BFGS-B', jac=self.grad, bounds=None, tol=1e-4,
                                callback=None, options={'maxcor': 10, 'maxfun': 15000, 'maxiter': 15000,
                                                        'maxls': 20})
        self.int_wpts = np.reshape(res.x[:nq], (self.D, self.M - 1))
        self.tau = res.x[nq:]
        self.ts = self.tau2T(self.tau)
        self.iter_num += res.nit
        self.nfev += res.nfev
        self.opt_running_times += 1
        self.last_status = res.message
        self.last_nit = res.nit
        self.weighted_cost = self.costs * self.p.weights
        self.final_cost = self.weighted_cost.sum()
        if self.weighted_cost[3] > self.p.collision_cost_tol:
            raise ValueError("collision cost too large")

    def warm_start_plan(self, grid, head_state, tail_state, int_wpts, ts):
        self.set_problem(grid, head_state, tail_state, int_wpts, ts)
        seed = 0
        while seed < 5:
            try:
                self.plan_once()
                self.attempt = seed
                return
            except Exception:
                seed += 1
                self.int_wpts, self.ts = self.straight_line_guess(head_state, tail_state, seed)
        raise Exception("No solution for the given target")

    def plan(self, grid, head_state, tail_state):
        w, ts = self.straight_line_guess(head_state, tail_state)
        self.warm_start_plan(grid, head_state, tail_state, w, ts)

    def batch_plan(self, grid, head_state, tail_state):
        cands, ts = self.lateral_guesses(head_state, tail_state)
        k = cands.shape[0]
        best_w = np.zeros(cands.shape)
        best_t = np.zeros((k, len(ts)))
        score = np.zeros(k)
        for i in range(k):
            try:
                self.set_problem(grid, head_state, tail_state, cands[i], ts)
                self.plan_once()
                best_w[i] = self.int_wpts
                best_t[i] = self.ts
                score[i] = self.weighted_cost.sum()
            except Exception:
                score[i] = np.inf
            if np.min(score) < np.inf:
                j = np.argmin(score)
                self.int_wpts = best_w[j]
                self.ts = best_t[j]
                self.final_cost = score[j]
                self.best_idx = int(j)
            else:
                self.warm_start_plan(grid, head_state, tail_state, cands[0], ts)
                self.best_idx = -1

    # ---- trajectory evaluation (TU:85-195) ------------------------------------------------
    def final_coeffs(self):
        A, rhs = minco_system(self.ts, self.head, self.tail, self.int_wpts)
        self.A = A
        self.coeffs = np.linalg.solve(A, rhs)
        return self.coeffs

    def _locate(self, t):
        k = 0
        while sum(self.ts[:k+1]) < t:
            k += 1
        return k, t - sum(self.ts[:k])

    def state_at(self, t, order):
        if t > sum(self.ts):
            return self.state_at(sum(self.ts), order)
        k, T = self._locate(t)
        blk = self.coeffs[6*k:6*(k+1), :]
        if order == 0:
            beta = np.array([1, T, T**2, T**3, T**4, T**5])
        elif order == 1:
            beta = np.array([0, 1, 2*T, 3*T**2, 4*T**3, 5*T**4])
        elif order == 2:
            beta = np.array([0, 0, 2, 6*T, 12*T**2, 20*T**3])
        else:
            beta = np.array([0, 0, 0, 6, 24*T, 60*T**2])
        return np.dot(blk.T, np.array([beta]).T).T

    def full_state_cmd(self, hz=300):
        self.final_coeffs()
        tt = np.arange(0, sum(self.ts), 1/hz)
        out = np.zeros((tt.shape[0], 3, self.D))
        for i, t in enumerate(tt):
            for o in range(3):
                out[i][o] = self.state_at(t, o)
        return out

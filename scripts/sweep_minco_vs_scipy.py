"""Offline sweep (CPU): scipy.optimize.minimize against the restated L-BFGS-B (oracle/minco_oracle.c) on MINCO
problems of M pieces, both driven by the reference-identical Python evaluator (oracle/minco_ref.py): final x bit for
bit, nit, nfev.      python scripts/sweep_minco_vs_scipy.py <M> <count> [world]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.optimize as sopt
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig
from oracle import c_oracle, minco_ref
M = int(sys.argv[1]); count = int(sys.argv[2]); wid = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = YamlConfig(); cfg.init_wpts_num = M - 1
w = make_world(wid, dense=(M >= 8))
grid = minco_ref.GridMap(w.occ, w.H, w.W, w.res, w.ox, w.oy)
head, tail = make_problems(w, count, M=M)
opt = minco_ref.RefOptimizer(cfg)
same = ran = 0
for k in range(count):
    q0, ts0 = opt.straight_line_guess(head[k], tail[k])
    opt.set_problem(grid, head[k], tail[k], q0, ts0)
    x0 = np.concatenate((q0.reshape(-1), opt.T2tau(ts0)))
    try:
        res = sopt.minimize(opt.cost, x0, method='L-BFGS-B', jac=opt.grad, bounds=None, tol=1e-4,
                            options={'maxcor': 10, 'maxfun': 15000, 'maxiter': 15000, 'maxls': 20})
    except (OverflowError, ValueError, ZeroDivisionError):
        continue
    x, nit, nfev, st = c_oracle.lbfgsb_cb(lambda v: (opt.cost(v), opt.grad(v)), x0)
    ran += 1
    ok = bool(np.array_equal(x, res.x) and nit == res.nit and nfev == res.nfev)
    same += ok
    if not ok: print('MISMATCH problem', k, 'nit', nit, res.nit, 'nfev', nfev, res.nfev, flush=True)
print(f'M={M} (n={3*M-2}), world {wid}: {same}/{ran} minimize() runs bit-identical to scipy')

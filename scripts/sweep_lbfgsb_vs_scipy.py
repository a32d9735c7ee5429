"""Offline sweep (CPU): the restated L-BFGS-B (oracle/minco_oracle.c) against scipy.optimize.minimize on N random
objective functions of dimension 2..28 (the generator of tests/test_lbfgsb_property.py): final x bit for bit, nit, nfev.
    python scripts/sweep_lbfgsb_vs_scipy.py <seed> <N>        # round 2: 4 x 1500 cases, 0 mismatches"""
import numpy as np, scipy.optimize as sopt, sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle
bad=0; tot=0; t0=time.time()
rs=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
N=int(sys.argv[2]) if len(sys.argv)>2 else 2000
for it in range(N):
    n=int(rs.integers(2,29)); seed=int(rs.integers(0,2**31-1)); cond=float(rs.uniform(0,4)); quartic=float(rs.uniform(0,2)); skew=float(rs.choice([0.0,0.0,rs.uniform(0,0.3)]))
    rng = np.random.default_rng(seed)
    Q = np.linalg.qr(rng.normal(size=(n, n)))[0]
    H = Q @ np.diag(10.0 ** rng.uniform(-cond / 2, cond / 2, n)) @ Q.T
    c = rng.normal(size=n); S = rng.normal(size=(n, n)) * skew
    x0 = rng.normal(size=n) * 3.0
    f=lambda x: float(0.5 * x @ H @ x + quartic * np.sum((x - c) ** 4) + np.sum(np.cos(x)))
    g=lambda x: H @ x + 4.0 * quartic * (x - c) ** 3 - np.sin(x) + S @ np.sin(x)
    res = sopt.minimize(f, x0, method='L-BFGS-B', jac=g, bounds=None, tol=1e-4, options={'maxcor': 10, 'maxfun': 15000, 'maxiter': 15000, 'maxls': 20})
    x, nit, nfev, status = c_oracle.lbfgsb_cb(lambda v: (f(v), g(v)), x0)
    tot+=1
    if not (nit==res.nit and nfev==res.nfev and np.array_equal(x,res.x)):
        bad+=1; print('MISMATCH', n, seed, cond, quartic, skew, nit, res.nit, nfev, res.nfev, flush=True)
print('cases', tot, 'mismatches', bad, 'seconds', round(time.time()-t0,1))

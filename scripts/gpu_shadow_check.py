"""EXPERIMENTAL speculative restarts (NEO_SHADOW=1; lbfgs_warp.cuh, protocol = oracle/shadow_sim.c) -- first hardware check,
to be run under a shell timeout (the path has not run on a GPU yet):

    timeout 120 python scripts/gpu_shadow_check.py [n_problems]

Runs the bench workload (M = 3, 5 attempts) with NEO_SHADOW=0 and =1 on identical inputs and prints both kernel times and
whether every output (x, ts, coeffs, costs, status, ok, attempt, nit, runs, nfev) is bit-identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from neo_planner_b200 import lib, guesses  # noqa: E402
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = 3
cfg = YamlConfig()
w = make_world(0)
head, tail = make_problems(w, B, M=M)
q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(1))
res = {}
for v in ('0', '1'):
    os.environ['NEO_SHADOW'] = v
    h = lib.Handle(cfg, 0, 1)
    h.set_map_occupancy(0, w.H, w.W, w.res, w.ox, w.oy, w.occ)
    best = 1e9
    for _ in range(5):
        out = h.optimize(M, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=5)
        best = min(best, h.last_kernel_ms())
    res[v] = (best, {k: np.array(a) for k, a in out.items() if a is not None and k != 'work'})
    h.close()
    print(f'NEO_SHADOW={v}: kernel {best:.3f} ms, ok {out["ok"].mean():.3f}, mean evals {out["nfev"].mean():.1f}', flush=True)
diff = [k for k in res['0'][1] if not np.array_equal(res['0'][1][k], res['1'][1][k])]
print('identical outputs' if not diff else f'DIFFERENT: {diff}', f'speed-up {res["0"][0] / res["1"][0]:.2f}x')

"""A/B of kernel instantiations on identical inputs (development probe, run under gpurun):
    python scripts/gpu_ab_min_ctas.py ctas     # 2 vs 3 CTAs per SM (NEO_MIN_CTAS_FORCE)
    python scripts/gpu_ab_min_ctas.py staged   # packed by-piece kernel: all pieces parked vs staged per piece (NEO_STAGED)
Prints the best-of-4 kernel time of each variant and whether the results are bit-identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from neo_planner_b200 import lib, guesses  # noqa: E402
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'ctas'
if what == 'ctas':
    var, values = 'NEO_MIN_CTAS_FORCE', ('2', '3')
    cases = [(3, 2048), (3, 4096), (3, 8192), (3, 32768), (10, 4096), (10, 16384)]
else:
    var, values = 'NEO_STAGED', ('0', '1')
    cases = [(3, 16384), (3, 32768), (3, 65536), (4, 16384), (2, 32768)]
for M, B in cases:
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    w = make_world(0, dense=(M == 10))
    head, tail = make_problems(w, B, M=M)
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(1))
    res = {}
    for v in values:
        os.environ[var] = v
        h = lib.Handle(cfg, 0, 1)
        h.set_map_occupancy(0, w.H, w.W, w.res, w.ox, w.oy, w.occ)
        best = 1e9
        for _ in range(4):
            out = h.optimize(M, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=5)
            best = min(best, h.last_kernel_ms())
        res[v] = (best, out['x'].copy(), out['costs'].copy(), out['nfev'].copy())
        h.close()
    a, b = res[values[0]], res[values[1]]
    same = np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    print(f'M={M} B={B}: {var}={values[0]} {a[0]:.3f} ms, ={values[1]} {b[0]:.3f} ms, ratio {a[0] / b[0]:.3f}, '
          f'identical results {same}', flush=True)

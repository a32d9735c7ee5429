import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib, guesses
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig
cases = [(3, 2048), (3, 4096), (3, 8192), (3, 32768), (10, 4096), (10, 16384)]
for M, B in cases:
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    w = make_world(0, dense=(M == 10))
    head, tail = make_problems(w, B, M=M)
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(1))
    res = {}
    for force in ('2', '3'):
        os.environ['NEO_MIN_CTAS_FORCE'] = force
        h = lib.Handle(cfg, 0, 1)
        h.set_map_occupancy(0, w.H, w.W, w.res, w.ox, w.oy, w.occ)
        best = 1e9
        for _ in range(4):
            out = h.optimize(M, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=5)
            best = min(best, h.last_kernel_ms())
        res[force] = (best, out['x'].copy())
        h.close()
    same = np.array_equal(res['2'][1], res['3'][1])
    print(f'M={M} B={B}: 2 CTAs/SM {res["2"][0]:.3f} ms, 3 CTAs/SM {res["3"][0]:.3f} ms, ratio {res["2"][0]/res["3"][0]:.3f}, identical results {same}', flush=True)

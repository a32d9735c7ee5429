"""Critical-path probe: which problems occupy a warp longest, and how long one evaluation takes inside the optimizer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib, guesses
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig
M, B = 3, 1024
cfg = YamlConfig()
w = make_world(0)
head, tail = make_problems(w, B, M=M)
q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(1))
h = lib.Handle(cfg, 0, 1)
h.set_map_occupancy(0, w.H, w.W, w.res, w.ox, w.oy, w.occ)
for att in (1, 5):
    for _ in range(3):
        out = h.optimize(M, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=att)
    ms = h.last_kernel_ms()
    ns = out['work'][:, 3]
    k = np.argsort(-ns)[:5]
    print(f'attempts={att}: kernel {ms:.3f} ms; slowest problems (us on a warp, nfev, nit, attempt): '
          + ', '.join(f'({ns[i]/1e3:.0f}, {out["nfev"][i]}, {out["nit"][i]}, {out["attempt"][i]})' for i in k))
    tot = out['nfev'].sum()
    print(f'   sum warp-time {ns.sum()/1e6:.2f} ms over {tot} evals -> {ns.sum()/tot/1e3:.2f} us per eval (optimizer included); '
          f'slowest: {ns[k[0]]/out["nfev"][k[0]]/1e3:.2f} us per eval')

"""Development: kernel time of a batch of four-piece problems (M = 4: two problems per warp from 12,288 problems)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib, guesses
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig
M = 4; B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cfg = YamlConfig(); cfg.init_wpts_num = M - 1
ws = [make_world(20 + k) for k in range(16)]
per = B // len(ws)
heads, tails, ids = [], [], []
for k, w in enumerate(ws):
    a, b = make_problems(w, per, M=M); heads.append(a); tails.append(b); ids.append(np.full(per, k, np.int32))
head, tail, ids = np.concatenate(heads), np.concatenate(tails), np.concatenate(ids)
q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(1))
h = lib.Handle(cfg, 0, len(ws))
for k, w in enumerate(ws):
    h.set_map_occupancy(k, w.H, w.W, w.res, w.ox, w.oy, w.occ)
best = 1e9
for rep in range(3):
    out = h.optimize(M, q0, ts0, lib.pad_state(head), lib.pad_state(tail), ids, rq, rts, 5)
    best = min(best, h.last_kernel_ms())
print(B, 'problems, M = 4:', round(best, 3), 'ms, ok', out['ok'].mean(), 'mean evals', out['nfev'].mean())

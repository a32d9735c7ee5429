"""Profiling target (run under ncu via gpurun): a few k_eval and k_optimize launches on workload c2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib, guesses
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig
M = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = YamlConfig(); cfg.init_wpts_num = M - 1
w = make_world(0, dense=(M == 10))
head, tail = make_problems(w, B, M=M)
q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(1))
h = lib.Handle(cfg, 0, 1)
h.set_map_occupancy(0, w.H, w.W, w.res, w.ox, w.oy, w.occ)
tau, _ = h.T2tau(ts0)
x = np.concatenate([q0.reshape(B, -1), tau], axis=1)
for _ in range(3):
    h.eval(M, x, head, tail)
for _ in range(3):
    out = h.optimize(M, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=5)
print('done', h.last_kernel_ms())

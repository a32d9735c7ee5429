"""Throughput probe for large batches (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib, guesses
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig
for M, B in [(3, 1024), (3, 8192), (3, 32768), (10, 16384)]:
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    w = make_world(0, dense=(M == 10))
    head, tail = make_problems(w, B, M=M)
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(1))
    h = lib.Handle(cfg, 0, 1)
    h.set_map_occupancy(0, w.H, w.W, w.res, w.ox, w.oy, w.occ)
    for _ in range(3):
        out = h.optimize(M, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=5)
    ms = h.last_kernel_ms()
    print(f'{os.environ.get("NEO_SO","default")[-12:]} M={M} B={B}: {ms:.3f} ms -> {B/ms*1e3:.0f} traj/s; evals/s {out["nfev"].sum()/ms*1e3:.3e}; ok {out["ok"].mean():.3f}')

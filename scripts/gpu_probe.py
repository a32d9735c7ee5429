"""Development probe (run under gpurun): latency of one fused evaluation, evaluation-count distribution, kernel time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib, guesses
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig

for M, B in [(3, 1024), (3, 8192), (10, 1024), (10, 16384)]:
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
    ev_ms = h.last_kernel_ms()
    for att in (1, 5):
        for _ in range(3):
            out = h.optimize(M, q0, ts0, head, tail, retry_q=rq, retry_ts=rts, max_attempts=att)
        ms = h.last_kernel_ms()
        nf = out['nfev']
        print(f'M={M} B={B} attempts={att}: eval kernel {ev_ms:.3f} ms; optimize kernel {ms:.3f} ms -> {B/ms*1e3:.0f} traj/s; '
              f'nfev mean {nf.mean():.1f} p50 {np.percentile(nf,50):.0f} p99 {np.percentile(nf,99):.0f} max {nf.max()}; '
              f'ok {out["ok"].mean():.3f}; sum nfev {nf.sum()}; us per eval on critical path {ms*1e3/nf.max():.2f}')
print('fp64 peak', h.fp64_peak())

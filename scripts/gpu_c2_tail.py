"""Development: where does the time of a small batch (config 2, one problem per warp) go? Per-problem evaluation counts,
attempts and on-device task times against the kernel time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib
from bench import workload
wl = workload('c2', 0, 1); B = wl['B']; M = wl['M']
h = lib.Handle(wl['cfg'], 0, 1)
w_ = wl['worlds'][0]
h.set_map_occupancy(0, w_.H, w_.W, w_.res, w_.ox, w_.oy, w_.occ)
hp, tp = lib.pad_state(wl['head']), lib.pad_state(wl['tail'])
for rep in range(3):
    out = h.optimize(M, wl['q0'], wl['ts0'], hp, tp, None, wl['retry_q'], wl['retry_ts'], 5)
print('kernel ms', h.last_kernel_ms())
nf = out['nfev']; at = out['attempt']; us = out['work'][:, 3] / 1e3
print('attempt histogram', np.bincount(at, minlength=5), 'ok', out['ok'].mean())
for name, v in (('nfev (summed over attempts up to the returned one)', nf), ('task microseconds (summed)', us)):
    print(name, 'mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f' % (v.mean(), np.percentile(v, 50), np.percentile(v, 90), np.percentile(v, 99), v.max()))
print('microseconds per evaluation: median %.2f' % np.median(us / np.maximum(nf, 1)))
worst = np.argsort(-us)[:5]
print('slowest problems:', [(int(i), int(nf[i]), int(at[i]), round(float(us[i]), 1)) for i in worst])

"""Host-side phases of neo_optimize on config 4 (development): NEO_HOST_TIMING=1 python scripts/gpu_e2e_phases.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib
from bench import workload
wl = workload('c4', 0, 1); B = wl['B']; M = wl['M']
h = lib.Handle(wl['cfg'], 0, len(wl['worlds']))
ws = wl['worlds']
h.set_maps_occupancy(np.arange(len(ws)), ws[0].H, ws[0].W, ws[0].res, [w_.ox for w_ in ws], [w_.oy for w_ in ws],
                     np.stack([np.asarray(w_.occ).reshape(w_.H, w_.W) for w_ in ws]))
hp, tp = lib.pad_state(wl['head']), lib.pad_state(wl['tail'])
out = lib.Handle.alloc_result(B, M)
for rep in range(6):
    t0 = time.perf_counter()
    h.optimize(M, wl['q0'], wl['ts0'], hp, tp, wl['map_ids'], wl['retry_q'], wl['retry_ts'], 5, out=out)
    print('python call %.3f ms' % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)

"""One warm-up + one profiled k_optimize launch on a bench workload (ncu target).
    [NEO_TILE=8] python scripts/gpu_profile_opt.py c4 16384"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neo_planner_b200 import lib
from bench import workload
name = sys.argv[1] if len(sys.argv) > 1 else 'c4'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
wl = workload(name, 0, 1)
sl = slice(0, B)
h = lib.Handle(wl['cfg'], 0, len(wl['worlds']))
for slot, w_ in enumerate(wl['worlds']):
    h.set_map_occupancy(slot, w_.H, w_.W, w_.res, w_.ox, w_.oy, w_.occ)
ids = None if wl['map_ids'] is None else wl['map_ids'][sl]
hp, tp = lib.pad_state(wl['head'][sl]), lib.pad_state(wl['tail'][sl])
for _ in range(2):
    out = h.optimize(wl['M'], wl['q0'][sl], wl['ts0'][sl], hp, tp, ids, wl['retry_q'][sl], wl['retry_ts'], 5)
    print(h.last_kernel_ms(), out['ok'].mean())

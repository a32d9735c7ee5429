"""A/B on the dense-map workload (c5): kernel ms. NEO_SO selects the build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neo_planner_b200 import lib
from bench import workload
wl = workload('c5', 0, 1)
h = lib.Handle(wl['cfg'], 0, 1)
w_ = wl['world']; h.set_map_occupancy(0, w_.H, w_.W, w_.res, w_.ox, w_.oy, w_.occ)
hp, tp = lib.pad_state(wl['head']), lib.pad_state(wl['tail'])
ms = []
for rep in range(4):
    out = h.optimize(wl['M'], wl['q0'], wl['ts0'], hp, tp, None, wl['retry_q'], wl['retry_ts'], 5)
    ms.append(h.last_kernel_ms())
print(os.environ.get('NEO_SO', 'default'), 'c5 ms', min(ms[1:]), 'ok', out['ok'].mean())

"""Timing of the geometric initializer on the GPU (k_astar) and of geo-initialised planning next to the straight-line
guess. Usage: python scripts/gpu_geo_probe.py [n_problems] -> one JSON line (also gpurun_out/geo_probe.json)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neo_planner_b200.geo import BatchGeoPlanner  # noqa: E402
from neo_planner_b200.worlds import make_problems, make_world, YamlConfig  # noqa: E402
from oracle import astar_ref, minco_ref  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
w = make_world(0)
bp = BatchGeoPlanner(YamlConfig(), max_maps=1)
bp.set_map(w)
out = {}
for name, Mlen in (('5m', 3), ('16.7m', 10)):
    head, tail = make_problems(w, n, M=Mlen)
    bp.handle.astar(head[:, 0], tail[:, 0])                 # warm-up (allocates the scratch)
    ms = []
    for _ in range(5):
        r = bp.handle.astar(head[:, 0], tail[:, 0])
        ms.append(bp.handle.last_kernel_ms())
    t0 = time.perf_counter()
    r = bp.handle.astar(head[:, 0], tail[:, 0])
    wall = (time.perf_counter() - t0) * 1e3
    # CPU checker on a bounded sample (heap formulation; the reference's own dict scan is slower still)
    gm = minco_ref.GridMap(w.occ, w.H, w.W, w.res, w.ox, w.oy)
    k = 16
    t0 = time.perf_counter()
    same = 0
    for i in range(k):
        path, found, nclosed = astar_ref.astar(gm, head[i, 0], tail[i, 0])
        four, _, _ = astar_ref.prune(gm, path)
        same += int(np.array_equal(np.array(four), r['pruned'][i]) and nclosed == r['closed'][i])
    cpu = (time.perf_counter() - t0) / k
    out[name] = dict(problems=n, kernel_ms=float(np.median(ms)), e2e_ms=wall, paths_per_s=n / (np.median(ms) * 1e-3),
                     mean_closed=float(r['closed'].mean()), max_closed=int(r['closed'].max()), mean_path_len=float(r['path_len'].mean()),
                     cpu_oracle_ms_per_path=cpu * 1e3, oracle_agree=f'{same}/{k}')
head, tail = make_problems(w, n, M=3)
rng = np.random.default_rng(0)
a = bp.plan(head, tail, rng=np.random.default_rng(0))
g = bp.geo_plan(head, tail, rng=np.random.default_rng(0))
out['plan_5m'] = dict(straight=dict(ok=float(a['ok'].mean()), nit=float(a['nit'].mean()), nfev=float(a['nfev'].mean()), runs=float(a['runs'].mean())),
                      geo=dict(ok=float(g['ok'].mean()), nit=float(g['nit'].mean()), nfev=float(g['nfev'].mean()), runs=float(g['runs'].mean())))
print(json.dumps(out))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/geo_probe.json', 'w'), indent=1)

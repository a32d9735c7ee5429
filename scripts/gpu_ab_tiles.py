"""A/B of the lanes-per-problem choice (NEO_TILE) on the bench workloads: kernel ms per call, traj/s.
    python scripts/gpu_ab_tiles.py [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib
from bench import workload

res = {}
for name, sizes in ((('c2', [1024]), ('c4', [4096, 16384, 65536])) if not os.environ.get('AB_C4_ONLY') else (('c4', [16384, 65536]),)):
    wl_full = workload(name, 0, 1)
    for B in sizes:
        sl = slice(0, B)
        for tile in (32, 8):
            os.environ['NEO_TILE'] = str(tile)
            wl = wl_full
            h = lib.Handle(wl['cfg'], 0, len(wl['worlds']))
            for slot, w_ in enumerate(wl['worlds']):
                h.set_map_occupancy(slot, w_.H, w_.W, w_.res, w_.ox, w_.oy, w_.occ)
            ids = None if wl['map_ids'] is None else wl['map_ids'][sl]
            hp, tp = lib.pad_state(wl['head'][sl]), lib.pad_state(wl['tail'][sl])
            out = None
            ms = []
            for rep in range(4):
                out = h.optimize(wl['M'], wl['q0'][sl], wl['ts0'][sl], hp, tp, ids, wl['retry_q'][sl], wl['retry_ts'], 5, out=out)
                ms.append(h.last_kernel_ms())
            best = min(ms[1:])
            res[f'{name}_B{B}_tile{tile}'] = dict(ms=best, traj_per_s=B / best * 1e3, ok=float(out['ok'].mean()), nfev=float(out['nfev'].mean()))
            print(name, B, tile, res[f'{name}_B{B}_tile{tile}'], flush=True)
            del h
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], 'w'), indent=1)

"""Development probe: cycles per optimizer phase (library built with -DNEO_FAST_BUILD -DNEO_OPT_TICKS).
    NEO_SO=neo_planner_b200/libneoopt_ticks.so [NEO_TILE=8] python scripts/gpu_opt_ticks.py c2 1024"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib
from bench import workload
name = sys.argv[1]; B = int(sys.argv[2])
wl = workload(name, 0, 1); sl = slice(0, B)
h = lib.Handle(wl['cfg'], 0, len(wl['worlds']))
for slot, w_ in enumerate(wl['worlds'][:max(1, B // 256 + 1)]):
    h.set_map_occupancy(slot, w_.H, w_.W, w_.res, w_.ox, w_.oy, w_.occ)
ids = None if wl['map_ids'] is None else wl['map_ids'][sl]
hp, tp = lib.pad_state(wl['head'][sl]), lib.pad_state(wl['tail'][sl])
buf = np.zeros(64, np.int64)
names = {0: 'gd + dcsrch_step', 1: 'x87 norm', 2: 'lb_update', 3: 'factor: assemble', 4: 'factor: potf2 #1', 5: 'factor: trsm', 6: 'factor: (2,2) block',
         7: 'factor: potf2 #2', 8: 'step: wv', 9: 'step: trsv^T', 10: 'step: trsv', 11: 'step: combine', 12: 'accept tests + y', 13: 'z, d', 14: 'gd + dcsrch_start',
         15: 'before direction', 16: 'eval_fg', 17: ' eval: times_from_tau', 18: ' eval: load + solve nodes', 19: ' eval: hermite', 20: ' eval: energy', 21: ' eval: sample loop + reduce',
         22: ' eval: adjoint h', 23: ' eval: K^T solve', 24: ' eval: G rows', 25: ' eval: grad_T'}
for rep in range(2):
    out = h.optimize(wl['M'], wl['q0'][sl], wl['ts0'][sl], hp, tp, ids, wl['retry_q'][sl], wl['retry_ts'], 5)
    h.lib.neo_test_opt_ticks(buf.ctypes.data_as(C.c_void_p))
print('kernel ms', h.last_kernel_ms(), 'evals', out['nfev'].sum())
tot = buf[:17].sum()
for i, nm in names.items():
    if buf[32 + i]:
        print(f'{nm:24s} {buf[i] / buf[32 + i]:9.0f} cycles/call x {buf[32 + i]:9d} calls = {100 * buf[i] / tot:5.1f} %')

"""Development probe (library built with -DNEO_FAST_BUILD -DNEO_ROUND_TRACE, run with NEO_SM_SYNC=1): timeline of the
12 warps of CTA 0 over the first 2048 rounds -- how long each warp evaluates / advances and how long it waits."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import lib
from bench import workload
wl = workload('c4', 0, 1); B = 65536; sl = slice(0, B)
h = lib.Handle(wl['cfg'], 0, len(wl['worlds']))
for slot, w_ in enumerate(wl['worlds']):
    h.set_map_occupancy(slot, w_.H, w_.W, w_.res, w_.ox, w_.oy, w_.occ)
hp, tp = lib.pad_state(wl['head'][sl]), lib.pad_state(wl['tail'][sl])
buf = np.zeros(12 * 2048 * 4, np.int64)
for rep in range(2):
    out = h.optimize(wl['M'], wl['q0'][sl], wl['ts0'][sl], hp, tp, wl['map_ids'][sl], wl['retry_q'][sl], wl['retry_ts'], 5)
h.lib.neo_test_round_trace(buf.ctypes.data_as(C.c_void_p))
print('kernel ms', h.last_kernel_ms())
t = buf.reshape(12, 2048, 4).astype(np.float64)
R = slice(20, 600)
arrive, go, ev, adv = t[:, R, 0], t[:, R, 1], t[:, R, 2], t[:, R, 3]
nxt = t[:, 21:601, 0]
work = nxt - go; wait = go - arrive
ev = np.where(ev == 0, np.nan, ev)
te = ev - go; to = adv - ev; tail = nxt - adv
print('per round (cycles): work mean %.0f  max-over-warps mean %.0f  wait mean %.0f' % (work.mean(), work.max(axis=0).mean(), wait.mean()))
for nm, x in (('evaluate', te), ('advance', to), ('task end', tail)):
    print(f'  {nm:9s} mean {np.nanmean(x):8.0f}  std over warps {np.nanmean(np.nanstd(x, axis=0)):8.0f}  max-mean over warps {np.nanmean(np.nanmax(x, axis=0) - np.nanmean(x, axis=0)):8.0f}  p10 {np.nanpercentile(x, 10):8.0f} p50 {np.nanpercentile(x, 50):8.0f} p90 {np.nanpercentile(x, 90):8.0f} p99 {np.nanpercentile(x, 99):8.0f}')
# which phase decides the slowest warp
slow = work.argmax(axis=0)
idx = np.arange(work.shape[1])
print('slowest warp of a round: evaluate %.0f advance %.0f task end %.0f (means, cycles)' % (np.nanmean(te[slow, idx]), np.nanmean(to[slow, idx]), np.nanmean(tail[slow, idx])))
np.save('gpurun_out/round_trace.npy', t)

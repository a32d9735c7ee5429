"""Offline analysis (CPU, C checker with its line-search trace) of the dependent evaluation chain that bounds the
latency-bound launch (bench workload c2: 1,024 problems, M = 3, <= 5 attempts run speculatively in parallel):
how long the longest chain is today and how long it would be if the restart after a failed line search were computed
speculatively. Usage: python scripts/analyze_chain.py [n_problems]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neo_planner_b200 import guesses  # noqa: E402
from neo_planner_b200.worlds import make_problems, make_world, YamlConfig  # noqa: E402
from oracle import c_oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = YamlConfig(); M = 3
w = make_world(0)
head, tail = make_problems(w, n, M=M)
q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(0))
p = c_oracle.Params.from_config(cfg)
m = c_oracle.OracleMap.from_world(w)
lib = c_oracle.lib()
buf = np.zeros(1 << 16, np.int32)
now, shadow, evals_total, failed_total, n_failed_ls, n_ls = [], [], 0, 0, 0, 0
for b in range(n):
    lib.orc_set_trace(buf.ctypes.data_as(C.c_void_p), C.c_int(buf.size))
    c_oracle.plan_batch(p, m, M, head[b:b + 1], tail[b:b + 1], q0[b:b + 1], ts0[b:b + 1], rq[b:b + 1], rts, 5)
    k = lib.orc_trace_count()
    tr = buf[:k]
    chains_now, chains_sh = [], []
    for v in tr:
        if v == -1:
            chains_now.append(1); chains_sh.append(1)          # the evaluation at x0
            continue
        ev, failed, mem = v >> 2, (v >> 1) & 1, v & 1
        chains_now[-1] += ev
        n_ls += 1
        if failed and mem:
            n_failed_ls += 1; failed_total += ev                # restart known in advance: free with a shadow warp
        else:
            chains_sh[-1] += ev
        evals_total += ev
    if chains_now:                                              # attempts run concurrently: the slowest one counts
        now.append(max(chains_now)); shadow.append(max(chains_sh))
lib.orc_set_trace(None, C.c_int(0))
now, shadow = np.array(now), np.array(shadow)
print(f'{n} problems: {n_ls} line searches, {n_failed_ls} failed with non-empty memory '
      f'({100 * failed_total / max(evals_total, 1):.1f} % of all evaluations)')
for name, a in (('today', now), ('with speculative restarts', shadow)):
    print(f'  longest chain per problem, {name}: max {a.max()}, p99 {np.percentile(a, 99):.0f}, p90 {np.percentile(a, 90):.0f}, '
          f'mean {a.mean():.1f} evaluations')
print(f'  launch time is bound by the max: {now.max()} -> {shadow.max()} evaluations ({shadow.max() / now.max():.2f}x)')

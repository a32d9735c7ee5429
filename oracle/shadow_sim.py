"""shadow_sim.py -- TEST INFRASTRUCTURE ONLY. ctypes wrapper of oracle/shadow_sim.c, the multi-threaded protocol simulator
for speculative restarts. `plan_batch` has the inputs of c_oracle.plan_batch plus a thread
count and returns the same outputs plus protocol statistics."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import c_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'libshadow_sim.so')
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, 'shadow_sim.c'), os.path.join(HERE, 'minco_oracle.c')]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.run(['gcc', '-O2', '-fPIC', '-std=c11', '-ffp-contract=off', '-fno-fast-math', '-pthread', '-shared', '-o', SO,
                        srcs[0], '-lm'], check=True, cwd=HERE)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def plan_batch(params, omap, M, head, tail, q0, ts0, retry_q, retry_ts, max_attempts, threads):
    f64, _p = c_oracle.f64, c_oracle._p
    q0 = f64(q0); B = q0.shape[0]; n = 2 * (M - 1) + M
    ts0 = f64(ts0); head = f64(c_oracle.pad_state(head)); tail = f64(c_oracle.pad_state(tail))
    retry_q = f64(retry_q); retry_ts = f64(retry_ts)
    out = dict(x=np.zeros((B, n)), ts=np.zeros((B, M)), costs=np.zeros((B, 4)), status=np.zeros(B, np.int32),
               ok=np.zeros(B, np.int32), attempt=np.zeros(B, np.int32), nit=np.zeros(B, np.int32),
               runs=np.zeros(B, np.int32), nfev=np.zeros(B, np.int32))
    stats = np.zeros(4, np.int64)
    rc = lib().sim_plan_batch(C.byref(params), C.byref(omap.c), C.c_int(B), C.c_int(M), _p(head), _p(tail), _p(q0), _p(ts0),
                              _p(retry_q), _p(retry_ts), C.c_int(max_attempts), C.c_int(threads), _p(out['x']), _p(out['ts']),
                              _p(out['costs']), _p(out['status']), _p(out['ok']), _p(out['attempt']), _p(out['nit']),
                              _p(out['runs']), _p(out['nfev']), _p(stats))
    if rc != 0:
        raise RuntimeError(f'shadow simulator failed ({rc})')
    out['stats'] = dict(claims=int(stats[0]), handoffs=int(stats[1]), cancelled=int(stats[2]), speculative_evals=int(stats[3]))
    return out

"""Design check for speculative restarts, CPU only: every line search of the C checker that
fails with a non-empty memory is replayed from the state captured at its start {x, g, f, nit, evaluation count at the
failure}; the replay must end in the same result bit for bit. Usage: python scripts/check_restart_payload.py"""
import ctypes as C, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from neo_planner_b200 import guesses
from neo_planner_b200.worlds import make_problems, make_world, YamlConfig
from oracle import c_oracle
lib = c_oracle.lib()
for M, n, dense in ((3, 1024, False), (10, 128, True)):
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    w = make_world(0, dense=dense)
    head, tail = make_problems(w, n, M=M)
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(0))
    p = c_oracle.Params.from_config(cfg); m = c_oracle.OracleMap.from_world(w)
    hp = c_oracle.pad_state(head); tp = c_oracle.pad_state(tail)
    tot = bad = 0
    for b in range(n):
        for a in range(3):
            q = q0[b] if a == 0 else rq[b, a - 1]
            ts = ts0[b] if a == 0 else rts
            tau = -np.log((cfg.T_max - cfg.T_min) / (ts - cfg.T_min) - 1)
            x0 = np.ascontiguousarray(np.concatenate((q.reshape(-1), tau)))
            chk = C.c_int(0)
            bad += lib.orc_check_restarts(C.byref(p), C.byref(m.c), C.c_int(M), hp[b].ctypes.data_as(C.c_void_p), tp[b].ctypes.data_as(C.c_void_p),
                                          x0.ctypes.data_as(C.c_void_p), C.byref(chk))
            tot += chk.value
    print(f'M={M}: {tot} failed line searches replayed from their restart state, {bad} do not reproduce the original result')

"""CPU test of the speculative-restart protocol planned for k_optimize: oracle/shadow_sim.c lets
threads play persistent warps -- attempt-major task queue, speculative retries, and restart states that running tasks
publish for idle workers to claim (slot states REQUESTED / CLAIMED / CONFIRMED, epochs, hand-off of counters). Whatever the
interleaving, the outputs must equal the sequential checker's warm_start_plan bit for bit."""
import numpy as np

from neo_planner_b200 import guesses
from neo_planner_b200.worlds import make_problems, make_world, YamlConfig
from oracle import c_oracle, shadow_sim


def test_speculative_restart_protocol_reproduces_sequential_results():
    cfg = YamlConfig(); M = 3
    w = make_world(0)
    N = 96
    head, tail = make_problems(w, N, M=M)
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(0))
    p = c_oracle.Params.from_config(cfg); m = c_oracle.OracleMap.from_world(w)
    ref = c_oracle.plan_batch(p, m, M, head, tail, q0, ts0, rq, rts, 5)
    claims = handoffs = 0
    for chunk, threads in ((4, 16), (8, 8), (32, 3), (1, 6)):
        for s in range(0, N, chunk):
            sl = slice(s, s + chunk)
            out = shadow_sim.plan_batch(p, m, M, head[sl], tail[sl], q0[sl], ts0[sl], rq[sl], rts, 5, threads)
            for k in ('x', 'ts', 'costs', 'status', 'ok', 'attempt', 'nit', 'runs', 'nfev'):
                assert np.array_equal(out[k], ref[k][sl]), (k, s, chunk, threads)
            claims += out['stats']['claims']; handoffs += out['stats']['handoffs']
    assert claims > 50 and handoffs > 10        # the speculative path was actually exercised

"""CPU tests of the host-side logic: initial guesses (EP:82-140) are bit-identical to the Python restatement of the
reference (oracle/minco_ref.py, itself pinned to the reference by gen_golden.py), and the C-ABI library loads
and exports every symbol include/neoopt.h declares (no compute calls without a GPU)."""
import os
import re

import numpy as np
import pytest

from neo_planner_b200 import guesses, lib
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig
from oracle import minco_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('M', [2, 3, 6, 10])
def test_straight_line_guess_bit_identical(M):
    cfg = YamlConfig(); cfg.init_wpts_num = M - 1
    w = make_world(2)
    head, tail = make_problems(w, 64, M=M)
    head[5, 0, 1] = tail[5, 0, 1] = 1.25          # motion exactly along x: numpy.linspace's step == 0 branch
    head[6, 0, 0] = tail[6, 0, 0] = 3.0
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    ref = minco_ref.RefOptimizer(cfg)
    for b in range(64):
        q, ts = ref.straight_line_guess(head[b], tail[b])
        assert np.array_equal(q0[b], q) and np.array_equal(ts0[b], ts), b


def test_retry_guesses_follow_global_rng_stream():
    cfg = YamlConfig(); M = 3
    w = make_world(2)
    head, tail = make_problems(w, 4)
    ref = minco_ref.RefOptimizer(cfg)
    for b in range(4):
        np.random.seed(42 + b)
        mine, rts = guesses.retry_guesses(cfg, head[b], tail[b], M, 4)
        np.random.seed(42 + b)
        for a in range(4):
            q, ts = ref.straight_line_guess(head[b], tail[b], seed=a + 1)
            assert np.array_equal(mine[a], q) and np.array_equal(rts, ts)


def test_lateral_guesses_bit_identical():
    cfg = YamlConfig(); M = 3
    w = make_world(2)
    head, tail = make_problems(w, 32)
    cands, ts = guesses.lateral_guesses(cfg, head, tail, M)
    ref = minco_ref.RefOptimizer(cfg)
    for b in range(32):
        c, t = ref.lateral_guesses(head[b], tail[b])
        assert np.array_equal(cands[b], c) and np.array_equal(ts, t)


def test_adaptive_piece_count():
    cfg = YamlConfig(); cfg.init_wpts_mode = 'adaptive'
    head = np.array([[0.0, 0.0], [0, 0]]); tail = np.array([[5.0, 0.0], [0, 0]])
    assert guesses.pieces_for(cfg, head, tail) == 3          # ceil(5/2 - 1) = 2 waypoints
    tail[0, 0] = 1.0
    assert guesses.pieces_for(cfg, head, tail) == 2          # max(.., 1) waypoint


def test_library_exports_every_declared_symbol():
    from neo_planner_b200 import build
    build.build()
    hdr = open(os.path.join(ROOT, 'include', 'neoopt.h')).read()
    declared = set(re.findall(r'^(?:int|const char \*)\s*\*?(neo_\w+)\(', hdr, flags=re.M))
    assert len(declared) >= 20
    l = lib.load()
    for name in declared:
        assert hasattr(l, name), name
    assert declared == set(lib.EXPORTS)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(lib.NeoError, match='no usable CUDA device'):
        lib.Handle(YamlConfig())


def test_exp_dd_host_build_is_correctly_rounded():
    """The double-double exp (csrc/dd_exp.h), compiled for the host from the same header the kernels use, against
    mpmath at 200 bits: correctly rounded on every sample; and equal to libm's exp at the structural points
    tau0 = map_T2tau(k * 0.1) where a 1-ulp difference would flip int(T/delta_t) (EP:401)."""
    import math
    import mpmath as mp
    mp.mp.prec = 200
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-20, 20, 3000), rng.uniform(-700, 700, 500)])
    y = lib.exp_host(x)
    for xi, yi in zip(x, y):
        assert float(mp.exp(mp.mpf(float(xi)))) == yi, xi
    for T_min, T_max in [(0.5, 5.0), (2.0, 20.0)]:
        ts = np.arange(int(T_min * 10) + 1, int(T_max * 10)) * 0.1
        for T in np.concatenate([ts, ts * 1.5]):
            if not (T_min < T < T_max):
                continue
            tau = -math.log((T_max - T_min) / (T - T_min) - 1)
            e1 = lib.exp_host(np.array([-tau]))[0]
            T1 = (T_max - T_min) / (1 + e1) + T_min
            T2 = (T_max - T_min) / (1 + math.exp(-tau)) + T_min
            assert T1 == T2, T

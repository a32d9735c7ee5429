"""Property test (hypothesis) of the fused cost+gradient kernel against the CPU checker: random pillar maps,
origins and resolutions, random boundary states (with accelerations), random waypoints and durations anywhere in
(T_min, T_max), both parameter sets, 2..10 pieces, and both lane layouts (8/16 and 32 lanes per problem).
north_star tolerance: cost and gradient within 1e-6 relative; measured here ~1e-13."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from neo_planner_b200 import lib
from neo_planner_b200.worlds import YamlConfig, LibraryDefaultConfig
from oracle import c_oracle

pytestmark = pytest.mark.gpu

_handles = {}


def handle(cfg_name, tile):
    key = (cfg_name, tile)
    if key not in _handles:
        old = os.environ.get('NEO_TILE')
        os.environ['NEO_TILE'] = str(tile)
        try:
            _handles[key] = lib.Handle(YamlConfig() if cfg_name == 'yaml' else LibraryDefaultConfig(), 0, 1)
        finally:
            if old is None:
                del os.environ['NEO_TILE']
            else:
                os.environ['NEO_TILE'] = old
    return _handles[key]


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow], derandomize=True, database=None)
@given(seed=st.integers(0, 2**31 - 1), M=st.integers(2, 10), cfg_name=st.sampled_from(['yaml', 'default']),
       tile=st.sampled_from([8, 32]), res=st.sampled_from([0.05, 0.1, 0.25]))
def test_eval_matches_checker_on_random_maps_and_states(seed, M, cfg_name, tile, res):
    rng = np.random.default_rng(seed)
    cfg = YamlConfig() if cfg_name == 'yaml' else LibraryDefaultConfig()
    H, W = int(rng.integers(40, 160)), int(rng.integers(40, 160))
    ox, oy = float(rng.uniform(-20, 5)), float(rng.uniform(-20, 5))
    occ = np.zeros((H, W), np.int8)
    for _ in range(int(rng.integers(0, 12))):
        r, c = int(rng.integers(0, H)), int(rng.integers(0, W))
        occ[r:r + int(rng.integers(1, 8)), c:c + int(rng.integers(1, 8))] = 100
    occ[rng.random((H, W)) < 0.002] = -1                                        # unknown cells are free (ESDF:23)
    h = handle(cfg_name, tile)
    h.set_map_occupancy(0, H, W, res, ox, oy, occ)
    m = c_oracle.OracleMap(occ, H, W, res, ox, oy)
    B = 16
    span = np.array([W * res, H * res])
    head = np.zeros((B, 3, 2)); tail = np.zeros((B, 3, 2))
    head[:, 0] = np.array([ox, oy]) + rng.uniform(-0.1, 1.1, (B, 2)) * span       # some states start outside the map
    tail[:, 0] = np.array([ox, oy]) + rng.uniform(-0.1, 1.1, (B, 2)) * span
    for s in (head, tail):
        s[:, 1] = rng.normal(0, 0.8, (B, 2)); s[:, 2] = rng.normal(0, 0.4, (B, 2))
    lam = np.linspace(0, 1, M + 1)[1:-1]
    q = head[:, None, 0, :] + lam[None, :, None] * (tail[:, 0] - head[:, 0])[:, None, :] + rng.normal(0, 0.4, (B, M - 1, 2))
    tau = rng.normal(0, 1.5, (B, M))
    x = np.concatenate([np.transpose(q, (0, 2, 1)).reshape(B, -1), tau], axis=1)
    ev = h.eval(M, x, head, tail)
    costs, grad, status = c_oracle.eval_batch(c_oracle.Params.from_config(cfg), m, M, head, tail, x)
    assert np.array_equal(ev['status'], status)
    ok = status == 0
    assert np.allclose(ev['costs'][ok], costs[ok], rtol=1e-6, atol=1e-9)
    scale = np.maximum(np.max(np.abs(grad[ok]), axis=1), 1e-9)
    assert (np.max(np.abs(ev['grad'][ok] - grad[ok]), axis=1) <= 1e-6 * scale).all()

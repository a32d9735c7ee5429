"""Two ranks (gloo rendezvous, both on cuda:0 so it runs on a 1-GPU box) solve a 4-world problem list through the real
device path; the gathered result must equal the single-process multi-slot result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from neo_planner_b200 import sharding
from neo_planner_b200.worlds import make_world, make_problems, YamlConfig

pytestmark = pytest.mark.gpu
N_WORLDS, PER_WORLD, M = 4, 48, 3


def problem_list():
    heads, tails, wid = [], [], []
    for w in range(N_WORLDS):
        a, b = make_problems(make_world(10 + w), PER_WORLD)
        heads.append(a); tails.append(b); wid.append(np.full(PER_WORLD, w))
    return np.concatenate(heads), np.concatenate(tails), np.concatenate(wid)


def make_solver(rank, world_size):
    from neo_planner_b200.planner import BatchPlanner
    worlds = sharding.shard_worlds(N_WORLDS, world_size, rank)
    bp = BatchPlanner(YamlConfig(), device=0, max_maps=len(worlds))
    for slot, w in enumerate(worlds):
        bp.set_map(make_world(10 + w), slot)

    def solve(head, tail, map_ids):
        return bp.plan(head, tail, map_ids, max_attempts=1)
    return solve


def _worker(rank, world_size, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    head, tail, wid = problem_list()
    res = sharding.ShardedPlanner(make_solver(rank, world_size), M, rank, world_size).plan(head, tail, wid)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_equal_single_process():
    head, tail, wid = problem_list()
    single = sharding.ShardedPlanner(make_solver(0, 1), M, 0, 1).plan(head, tail, wid)
    assert single['ok'].mean() > 0.7
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for r in (0, 1):
        for k in ('x', 'ts', 'coeffs', 'costs', 'status', 'ok', 'nit', 'runs', 'nfev'):
            assert np.array_equal(out[r][k], single[k]), (r, k)


def test_device_shard_equals_host_path():
    """sharding.DeviceShard (device-pointer entry, records packed and gathered on the device, one D2H) against
    Handle.optimize on the same inputs: every field identical. One rank (NCCL group of size 1 is not needed: the
    single-rank path copies the packed buffer directly)."""
    import torch
    from neo_planner_b200 import guesses, lib
    cfg = YamlConfig()
    worlds = [make_world(10 + w) for w in range(2)]
    heads, tails, ids = [], [], []
    for slot, w in enumerate(worlds):
        a, b = make_problems(w, 96)
        heads.append(a); tails.append(b); ids.append(np.full(96, slot, np.int32))
    head, tail, ids = np.concatenate(heads), np.concatenate(tails), np.concatenate(ids)
    B = len(ids)
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(3))
    h = lib.Handle(cfg, 0, len(worlds))
    for slot, w in enumerate(worlds):
        h.set_map_occupancy(slot, w.H, w.W, w.res, w.ox, w.oy, w.occ)
    ref = h.optimize(M, q0, ts0, lib.pad_state(head), lib.pad_state(tail), ids, rq, rts, 5)
    with torch.cuda.stream(torch.cuda.Stream()):
        shard = sharding.DeviceShard(h, M, B, 1, 'cuda:0', 5)
        for _ in range(2):                                                  # buffers are reused between calls
            out = shard.plan(q0, ts0, head, tail, ids, rq, rts)
    for k in ('x', 'ts', 'coeffs', 'costs', 'status', 'ok', 'attempt', 'nit', 'runs', 'nfev'):
        assert np.array_equal(out[k][0], ref[k]), k

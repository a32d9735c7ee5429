"""Multi-GPU sharding of independent planning problems: one process per GPU, problems partitioned
across ranks with no data-path collective, and ONE all-gather of fixed-size result records at the end of a batch
(torch.distributed; NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests). The reference has no distributed
code; this is the B200-native way to run its expert data-generation sweep (config 4) on an 8-GPU box."""
from __future__ import annotations

import numpy as np

INT_FIELDS = ('status', 'ok', 'attempt', 'nit', 'runs', 'nfev')


def shard_worlds(n_worlds: int, world_size: int, rank: int):
    """Contiguous block of world ids owned by `rank` (each GPU uploads only its own maps)."""
    per, extra = divmod(n_worlds, world_size)
    start = rank * per + min(rank, extra)
    return list(range(start, start + per + (1 if rank < extra else 0)))


def shard_problems(world_of_problem, world_size: int, rank: int):
    """Indices of the problems whose world belongs to `rank` under shard_worlds."""
    world_of_problem = np.asarray(world_of_problem)
    n_worlds = int(world_of_problem.max()) + 1 if world_of_problem.size else 0
    mine = np.zeros(max(n_worlds, 1), dtype=bool)
    mine[shard_worlds(n_worlds, world_size, rank)] = True
    return np.nonzero(mine[world_of_problem])[0]


def record_width(M: int) -> int:
    return (3 * M - 2) + M + 12 * M + 4 + len(INT_FIELDS)


def pack_records(res, M: int) -> np.ndarray:
    """dict of result arrays (neo_result) -> (B, record_width) float64 (integers are exact in fp64)."""
    B = res['x'].shape[0]
    rec = np.empty((B, record_width(M)))
    o = 0
    for key in ('x', 'ts', 'coeffs', 'costs'):
        a = res[key].reshape(B, -1)
        rec[:, o:o + a.shape[1]] = a
        o += a.shape[1]
    for key in INT_FIELDS:
        rec[:, o] = res[key]
        o += 1
    assert o == record_width(M)
    return rec


def unpack_records(rec: np.ndarray, M: int):
    n = 3 * M - 2
    B = rec.shape[0]
    out, o = {}, 0
    for key, w, shape in (('x', n, (B, n)), ('ts', M, (B, M)), ('coeffs', 12 * M, (B, 6 * M, 2)), ('costs', 4, (B, 4))):
        out[key] = rec[:, o:o + w].reshape(shape).copy(); o += w
    for key in INT_FIELDS:
        out[key] = rec[:, o].astype(np.int32); o += 1
    return out


_pinned = {}


def _staging(name, shape, torch, pin):
    """Reusable host staging tensors (pinned when a CUDA device is used) so that the copies can run asynchronously."""
    key = (name, tuple(shape), pin)
    t = _pinned.get(key)
    if t is None:
        t = torch.empty(shape, dtype=torch.float64, pin_memory=pin)
        _pinned[key] = t
    return t


def gather_blocks(local_rec: np.ndarray, counts, device=None):
    """All-gather for the usual partition: rank r owns a CONTIGUOUS block of counts[r] problems and the blocks follow
    each other in rank order (shard_problems on a world-sorted problem list). No index column, no scatter: the blocks
    are padded to the largest count, gathered with ONE collective and land in global order. Returns (sum(counts),
    width); with equal counts this is a view of the reusable (pinned) staging buffer -- valid until the next call."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size()
    assert len(counts) == ws and local_rec.shape[0] == counts[dist.get_rank()]
    width, cap = local_rec.shape[1], max(counts)
    dev = torch.device('cpu') if device is None else torch.device(device)
    pin = dev.type == 'cuda'
    send = _staging('send_blocks', (cap, width), torch, pin)
    send.numpy()[:local_rec.shape[0]] = local_rec
    buf = send.to(dev, non_blocking=True) if pin else send
    allbuf = torch.empty((ws * cap, width), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allbuf, buf)
    recv = _staging('recv_blocks', (ws * cap, width), torch, pin)
    recv.copy_(allbuf, non_blocking=pin)
    if pin:
        torch.cuda.current_stream(dev).synchronize()
    rows = recv.numpy()
    if all(c == cap for c in counts):
        return rows
    return np.concatenate([rows[r * cap:r * cap + c] for r, c in enumerate(counts)], axis=0)


def gather_records(local_rec: np.ndarray, local_idx: np.ndarray, total: int, device=None, counts=None, offset: int = 0):
    """All-gather the per-rank records into global problem order on every rank. Ranks may own different counts:
    records are padded to the maximum count (one collective of fixed size, then the padding is dropped).
    counts: per-rank record counts if the caller knows them (saves the object gather); offset is added to local_idx.
    On a CUDA device the records travel through pinned staging buffers (H2D, NCCL all-gather, D2H)."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size()
    width = local_rec.shape[1]
    if counts is None:
        counts = [None] * ws
        dist.all_gather_object(counts, int(local_rec.shape[0]))
    cap = max(counts)
    dev = torch.device('cpu') if device is None else torch.device(device)
    pin = dev.type == 'cuda'
    nloc = local_rec.shape[0]
    send = _staging('send', (cap, width + 1), torch, pin)
    if nloc:
        sv = send.numpy()
        sv[:nloc, :width] = local_rec
        sv[:nloc, width] = np.asarray(local_idx, dtype=np.float64) + float(offset)
    buf = send.to(dev, non_blocking=True) if pin else send.clone()
    allbuf = torch.empty((ws * cap, width + 1), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allbuf, buf)
    recv = _staging('recv', (ws * cap, width + 1), torch, pin)
    recv.copy_(allbuf, non_blocking=pin)
    if pin:
        torch.cuda.current_stream(dev).synchronize()
    rows_all = recv.numpy().reshape(ws, cap, width + 1)
    starts = np.concatenate(([0], np.cumsum(counts)))
    if all(c == 0 or (rows_all[r, 0, width] == starts[r] and rows_all[r, c - 1, width] == starts[r] + c - 1) for r, c in enumerate(counts)) \
            and starts[-1] == total:
        # ranks own contiguous blocks in rank order (the world-wise partition): no scatter, one copy per rank
        out = np.empty((total, width))
        for r, c in enumerate(counts):
            idx = rows_all[r, :c, width]
            if c and not np.array_equal(idx, np.arange(starts[r], starts[r] + c)):
                break
            out[starts[r]:starts[r] + c] = rows_all[r, :c, :width]
        else:
            return out
    out = np.zeros((total, width))
    seen = np.zeros(total, dtype=bool)
    for r in range(ws):
        rows = rows_all[r, :counts[r]]
        idx = rows[:, width].astype(np.int64)
        out[idx] = rows[:, :width]
        seen[idx] = True
    if not seen.all():
        raise RuntimeError(f'{int((~seen).sum())} problems were not solved by any rank')
    return out


class ShardedPlanner:
    """Runs `solve(head, tail, local_map_ids) -> result dict` on this rank's share of a global problem list and returns
    the gathered results (global order) on every rank. `solve` is normally BatchPlanner.plan bound to this rank's GPU;
    the CPU tests pass a stand-in so the partition/gather logic is covered without a device."""

    def __init__(self, solve, M: int, rank: int, world_size: int, device=None):
        self.solve, self.M, self.rank, self.world_size, self.device = solve, M, rank, world_size, device

    def local_worlds(self, n_worlds):
        return shard_worlds(n_worlds, self.world_size, self.rank)

    def plan(self, head, tail, world_of_problem):
        head = np.asarray(head); tail = np.asarray(tail); world_of_problem = np.asarray(world_of_problem)
        idx = shard_problems(world_of_problem, self.world_size, self.rank)
        n_worlds = int(world_of_problem.max()) + 1
        first = self.local_worlds(n_worlds)[0] if len(idx) else 0
        if len(idx):
            res = self.solve(head[idx], tail[idx], (world_of_problem[idx] - first).astype(np.int32))
            rec = pack_records(res, self.M)
        else:
            rec = np.zeros((0, record_width(self.M)))
        if self.world_size == 1:
            full = np.zeros((len(head), rec.shape[1])); full[idx] = rec
        else:
            # every rank knows the whole problem list, hence every rank's share: contiguous blocks in rank order (a
            # world-sorted list) take the scatter-free path
            shares = [shard_problems(world_of_problem, self.world_size, r) for r in range(self.world_size)]
            blocks = np.concatenate(shares)
            if np.array_equal(blocks, np.arange(len(head))):
                full = gather_blocks(rec, [len(s) for s in shares], self.device)
            else:
                full = gather_records(rec, idx, len(head), self.device, counts=[len(s) for s in shares])
        return unpack_records(full, self.M)

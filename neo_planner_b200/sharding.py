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


def packed_record_bytes(B: int, M: int) -> int:
    """Bytes of one rank's packed record buffer: [x | ts | coeffs | costs] doubles, then 6 int32 fields, B problems."""
    return B * ((3 * M - 2) + M + 12 * M + 4) * 8 + B * len(INT_FIELDS) * 4


def packed_record_offsets(B: int, M: int):
    """Byte offsets of the fields inside one rank's packed record buffer (the layout neo_optimize_dev is pointed at)."""
    n = 3 * M - 2
    d = np.cumsum([0, B * n, B * M, B * 12 * M, B * 4]) * 8
    off = dict(x=int(d[0]), ts=int(d[1]), coeffs=int(d[2]), costs=int(d[3]))
    for i, k in enumerate(INT_FIELDS):
        off[k] = int(d[4]) + 4 * B * i
    return off


def split_packed_records(raw: np.ndarray, B: int, M: int):
    """(world_size, packed_record_bytes) uint8 -> per-field views, rank-major: x (ws, B, n), ts (ws, B, M), coeffs
    (ws, B, 6M, 2), costs (ws, B, 4), and the int32 fields (ws, B). No copies."""
    ws = raw.shape[0]
    n = 3 * M - 2
    nd = B * (n + M + 12 * M + 4)
    dbl = raw[:, :nd * 8].view(np.float64)
    ints = raw[:, nd * 8:].view(np.int32).reshape(ws, len(INT_FIELDS), B)
    o = np.cumsum([0, B * n, B * M, B * 12 * M, B * 4])
    out = dict(x=dbl[:, o[0]:o[1]].reshape(ws, B, n), ts=dbl[:, o[1]:o[2]].reshape(ws, B, M),
               coeffs=dbl[:, o[2]:o[3]].reshape(ws, B, 6 * M, 2), costs=dbl[:, o[3]:o[4]].reshape(ws, B, 4))
    for i, k in enumerate(INT_FIELDS):
        out[k] = ints[:, i]
    return out


class DeviceShard:
    """One rank's equal share (B problems) of a batch, solved through the device-pointer entry point with the records
    gathered ON THE DEVICE: host inputs -> pinned staging -> HBM, neo_optimize_dev writes ONE packed byte record buffer
    ([x | ts | coeffs | costs] doubles, then [status ok attempt nit runs nfev] int32), ONE all_gather_into_tensor (NCCL
    over NVLink) collects the buffers of all ranks, ONE device-to-host copy lands the gathered batch in pinned memory.
    The host-pointer path (ShardedPlanner + gather_blocks) moves every record over PCIe three times; this one once.

    plan() returns per-field views of the pinned buffer, rank-major: out['x'] has shape (world_size, B, n) -- global
    order for contiguous equal shards -- valid until the next plan()."""

    def __init__(self, handle, M, B, world_size, device, max_attempts=5):
        import torch
        self.torch, self.h, self.M, self.B, self.ws, self.A = torch, handle, M, B, world_size, max_attempts
        self.dev = torch.device(device)
        n, nq = 3 * M - 2, 2 * (M - 1)
        self.n, self.nq = n, nq
        self.nbytes = packed_record_bytes(B, M)
        pin = lambda *shape, dt=torch.float64: torch.empty(shape, dtype=dt, pin_memory=True)
        self.h_x0, self.h_head, self.h_tail = pin(B, n), pin(B, 6), pin(B, 6)
        self.h_ids = pin(B, dt=torch.int32)
        self.h_rq = pin(B, max(max_attempts - 1, 1) * nq)
        self.h_rtau = pin(M)
        self.d_x0, self.d_head, self.d_tail = (torch.empty_like(t, device=self.dev) for t in (self.h_x0, self.h_head, self.h_tail))
        self.d_ids = torch.empty_like(self.h_ids, device=self.dev)
        self.d_rq = torch.empty_like(self.h_rq, device=self.dev)
        self.d_rtau = torch.empty_like(self.h_rtau, device=self.dev)
        self.d_out = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.dev)
        self.d_all = torch.zeros(world_size * self.nbytes, dtype=torch.uint8, device=self.dev)
        self.h_all = torch.empty(world_size * self.nbytes, dtype=torch.uint8, pin_memory=True)
        from . import lib
        base = self.d_out.data_ptr()
        res = lib.Result()
        for k, o in packed_record_offsets(B, M).items():
            setattr(res, k, base + o)
        res.work = None
        self.res = res

    def plan(self, q0, ts0, head, tail, map_ids=None, retry_q=None, retry_ts=None):
        import ctypes as C
        import torch.distributed as dist
        from . import lib
        torch, h, B, M, n, nq = self.torch, self.h, self.B, self.M, self.n, self.nq
        tau0, st0 = h.T2tau(np.asarray(ts0, dtype=np.float64).reshape(B, M))          # host libm, like the reference (EP:207-211)
        if st0.any():
            raise ValueError('initial durations outside (T_min, T_max): use the host-pointer path, which reports them per problem')
        x0 = self.h_x0.numpy()
        x0[:, :nq] = np.asarray(q0, dtype=np.float64).reshape(B, nq); x0[:, nq:] = tau0
        self.h_head.numpy()[:] = lib.pad_state(head).reshape(B, 6)
        self.h_tail.numpy()[:] = lib.pad_state(tail).reshape(B, 6)
        retry_status = 0
        if self.A > 1:
            self.h_rq.numpy()[:] = np.asarray(retry_q, dtype=np.float64).reshape(B, -1)
            rtau, rst = h.T2tau(np.asarray(retry_ts, dtype=np.float64))
            self.h_rtau.numpy()[:] = rtau
            retry_status = int(rst.max())
        if map_ids is not None:
            self.h_ids.numpy()[:] = map_ids
        for d, s in ((self.d_x0, self.h_x0), (self.d_head, self.h_head), (self.d_tail, self.h_tail), (self.d_rq, self.h_rq),
                     (self.d_rtau, self.h_rtau), (self.d_ids, self.h_ids)):
            d.copy_(s, non_blocking=True)
        st = torch.cuda.current_stream(self.dev).cuda_stream
        h._ck(h.lib.neo_optimize_dev(h.h, B, M, self.d_x0.data_ptr(), None, self.d_head.data_ptr(), self.d_tail.data_ptr(),
                                     None if map_ids is None else self.d_ids.data_ptr(),
                                     self.d_rq.data_ptr() if self.A > 1 else None, self.d_rtau.data_ptr() if self.A > 1 else None,
                                     retry_status, self.A, C.byref(self.res), C.c_void_p(st)))
        if self.ws > 1:
            dist.all_gather_into_tensor(self.d_all, self.d_out)
            self.h_all.copy_(self.d_all, non_blocking=True)
        else:
            self.h_all.copy_(self.d_out, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.views()

    def views(self):
        return split_packed_records(self.h_all.numpy().reshape(self.ws, self.nbytes), self.B, self.M)

"""Initial guesses of the reference planner, batched on the host (they are O(B*M) flops; the hot path is the
optimizer): generate_init_variables (EP:82-101) and batch_generate_init_variables (EP:103-140).

The arithmetic follows numpy.linspace(start + step, target, k, endpoint=False) operation by operation so the
batched result is bit-identical to calling the reference per problem (tests/test_host_logic.py)."""
from __future__ import annotations

import math

import numpy as np


def pieces_for(cfg, head, tail) -> int:
    """Number of pieces M the reference would use for one problem (EP:86-90)."""
    if cfg.init_wpts_mode == 'adaptive':
        length = np.linalg.norm(np.asarray(tail)[0] - np.asarray(head)[0])
        return max(math.ceil(length / cfg.init_seg_len - 1), 1) + 1
    return int(cfg.init_wpts_num) + 1


def _line(start, target, k):
    """Rows i = 0..k-1 of np.linspace(start + (target-start)/(k+1), target, k, endpoint=False) for
    start/target of shape (B, 2). Returns (B, k, 2)."""
    hop = (target - start) / (k + 1)
    s0 = start + hop
    delta = target - s0
    step = delta / k
    idx = np.arange(0, k, dtype=np.float64).reshape(1, k, 1)
    normal = idx * step[:, None, :] + s0[:, None, :]
    degenerate = (idx / k) * delta[:, None, :] + s0[:, None, :]       # numpy's branch when any(step == 0)
    zero = (step == 0).any(axis=1)
    return np.where(zero[:, None, None], degenerate, normal)


def init_ts(cfg, M):
    """EP:97-99."""
    ts = cfg.init_T * np.ones((M,))
    ts[0] *= 1.5
    ts[-1] *= 1.5
    return ts


def straight_line_guess(cfg, head, tail, M):
    """generate_init_variables(seed=0) for a batch. head/tail: (B, k, 2). Returns q0 (B, 2, M-1), ts0 (B, M)."""
    head = np.asarray(head, dtype=np.float64); tail = np.asarray(tail, dtype=np.float64)
    B = head.shape[0]
    w = _line(head[:, 0, :], tail[:, 0, :], M - 1)
    return np.ascontiguousarray(np.transpose(w, (0, 2, 1))), np.tile(init_ts(cfg, M), (B, 1))


def retry_guesses(cfg, head, tail, M, count, rng=None):
    """The `count` re-seeded guesses warm_start_plan would draw one after another (EP:200 -> EP:92-94):
    straight line + N(0, 0.5) noise of shape (M-1, 2) per retry.

    Single problem (head (k,2)): returns (count, 2, M-1) drawn from ``np.random.normal`` (the reference's
    global, unseeded stream) unless ``rng`` is given. Batch (head (B,k,2)): returns (B, count, 2, M-1); problem
    b consumes its `count` draws consecutively, i.e. the stream order of running the problems one by one with
    every retry taken."""
    head = np.asarray(head, dtype=np.float64); tail = np.asarray(tail, dtype=np.float64)
    single = head.ndim == 2
    if single:
        head = head[None]; tail = tail[None]
    B = head.shape[0]
    line = _line(head[:, 0, :], tail[:, 0, :], M - 1)                   # (B, M-1, 2)
    draw = np.random.normal if rng is None else rng.normal
    noise = draw(0, 0.5, (B, count, M - 1, 2))
    q = line[:, None, :, :] + noise
    q = np.ascontiguousarray(np.transpose(q, (0, 1, 3, 2)))
    return (q[0] if single else q), init_ts(cfg, M)


def lateral_guesses(cfg, head, tail, M, count=3, shift=0.6):
    """batch_generate_init_variables (EP:103-140) for a batch: candidate 0 is the straight line, the others are
    shifted +-0.6 m along the lateral normal. Returns (B, count, 2, M-1), ts (M)."""
    head = np.asarray(head, dtype=np.float64); tail = np.asarray(tail, dtype=np.float64)
    a, z = head[:, 0, :], tail[:, 0, :]
    diff = z - a
    nrm = np.array([np.linalg.norm(v) for v in diff])                 # per problem, the reference's call (EP:113)
    u = diff / nrm[:, None]
    side = np.stack([np.stack([u[:, 1], -u[:, 0]], axis=1), np.stack([-u[:, 1], u[:, 0]], axis=1)], axis=1)  # (B,2,2)
    line = _line(a, z, M - 1)                                          # (B, M-1, 2)
    cands = np.zeros((head.shape[0], count, M - 1, 2))
    cands[:, 0] = line
    flag = 0
    for j in range(1, count):
        cands[:, j] = line + shift * side[:, flag][:, None, :]
        flag = 1 - flag
    return np.ascontiguousarray(np.transpose(cands, (0, 1, 3, 2))), init_ts(cfg, M)

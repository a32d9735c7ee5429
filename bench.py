#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 MINCO trajectory optimizer (contract: see the task prompt / DESIGN.md §measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path (warm_start_plan semantics: straight-line expert guess, up to 5 L-BFGS-B
attempts, EP:62-80 / EP:186-237) over one batch of synthetic problems:
  workload c2 (default, BASELINE.json configs[1]): 1,024 start-goal pairs on one shared 300x300 @0.1 m random-pillar
  map, M = 3 pieces, planner_config.yaml parameters. With N ranks every rank gets its own world and its own 1,024
  problems (weak scaling, sharded by world as SURVEY.md §8e prescribes) and the packed result records are gathered
  with one NCCL all-gather per step.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from neo_planner_b200.worlds import make_world, make_problems, YamlConfig  # noqa: E402
from neo_planner_b200 import guesses  # noqa: E402

METRIC = 'optimized trajectories/sec to convergence'
UNIT = 'traj/s'


def workload(name, rank, world_size=1):
    """Seeded problem set of one rank. Returns dict(worlds, cfg, M, head, tail, map_ids, q0, ts0, retry_q, retry_ts).
    c2: 1,024 problems on one world per rank (weak scaling). c5: 16,384 problems, M = 10, 1200x1200 @0.05 m map per rank.
    c4: 65,536 problems = 256 worlds x 256 pairs in total, worlds sharded over the ranks (strong scaling)."""
    from neo_planner_b200 import sharding
    cfg = YamlConfig()
    if name == 'c5':
        M, per_world, dense, world_ids = 10, 16384, True, [rank]
    elif name == 'c4':
        M, per_world, dense, world_ids = 3, 256, False, sharding.shard_worlds(256, world_size, rank)
    else:
        M, per_world, dense, world_ids = 3, 1024, False, [rank]
    cfg.init_wpts_num = M - 1
    worlds, heads, tails, ids = [], [], [], []
    for slot, wid in enumerate(world_ids):
        w = make_world(wid, dense=dense)
        a, b = make_problems(w, per_world, M=M)
        worlds.append(w); heads.append(a); tails.append(b); ids.append(np.full(per_world, slot, np.int32))
    head, tail, map_ids = np.concatenate(heads), np.concatenate(tails), np.concatenate(ids)
    q0, ts0 = guesses.straight_line_guess(cfg, head, tail, M)
    rq, rts = guesses.retry_guesses(cfg, head, tail, M, 4, rng=np.random.default_rng(77_000 + rank))
    return dict(name=name, world=worlds[0], worlds=worlds, cfg=cfg, M=M, B=len(head), head=head, tail=tail,
                map_ids=map_ids if len(worlds) > 1 else None, q0=q0, ts0=ts0, retry_q=rq, retry_ts=rts,
                scaling='strong' if name == 'c4' else 'weak')


def describe(wl, n_gpus):
    w = wl['world']
    cfgidx = {'c2': 1, 'c4': 3, 'c5': 4}[wl['name']]
    maps = f"{len(wl['worlds'])} {w.H}x{w.W} @{w.res} m random-pillar map(s) per GPU"
    return {'workload': f"{wl['B']} start-goal pairs per GPU on {maps}, M={wl['M']} pieces, expert straight-line init, "
                        f"warm_start_plan semantics (<=5 attempts), planner_config.yaml parameters (BASELINE.json configs[{cfgidx}])",
            'problems_per_gpu': wl['B'], 'pieces': wl['M'], 'map': f'{w.H}x{w.W}@{w.res}', 'maps_per_gpu': len(wl['worlds']),
            'max_attempts': 5,
            'sharding': f'{n_gpus} rank(s), worlds sharded over ranks, NCCL all-gather of result records' if n_gpus > 1 else 'single GPU',
            'l2': 'flushed between timed steps (256 MiB write)'}


# ------------------------------------------------------------------------------------------------ CPU arms
_W = {}


def cpu_sample_workload(name):
    """The CPU arms time a bounded sample: the problems of the first world of the workload."""
    return workload(name, 0, 256 if name == 'c4' else 1)


def _pool_init(name, rank):
    import warnings
    warnings.filterwarnings('ignore')
    from oracle import minco_ref
    wl = cpu_sample_workload(name)
    w = wl['world']
    _W['wl'] = wl
    _W['grid'] = minco_ref.GridMap(w.occ, w.H, w.W, w.res, w.ox, w.oy)
    _W['opt'] = minco_ref.RefOptimizer(wl['cfg'])


def _pool_plan(k):
    """One reference-equivalent plan() (Python/NumPy + scipy L-BFGS-B, oracle/minco_ref.py) for problem k."""
    wl, opt = _W['wl'], _W['opt']
    np.random.seed(1_000 + k)
    n0, f0 = opt.iter_num, opt.nfev
    try:
        opt.plan(_W['grid'], wl['head'][k], wl['tail'][k])
        ok = 1
    except Exception:
        ok = 0
    return ok, opt.nfev - f0


def cpu_reference_run(name, steps, warmup, per_step=None):
    """Times the reference's own algorithm (Python + scipy, all host cores) on bounded samples of the workload."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_step = per_step or max(16, 2 * cores)
    ctx = mp.get_context('fork')
    with ctx.Pool(cores, initializer=_pool_init, initargs=(name, 0)) as pool:
        B = {'c2': 1024, 'c4': 256, 'c5': 16384}[name]
        k0 = 0
        for _ in range(warmup):
            pool.map(_pool_plan, [(k0 + i) % B for i in range(per_step)], chunksize=1)
            k0 += per_step
        t0 = time.perf_counter()
        oks = evals = 0
        for _ in range(steps):
            r = pool.map(_pool_plan, [(k0 + i) % B for i in range(per_step)], chunksize=1)
            k0 += per_step
            oks += sum(a for a, _ in r); evals += sum(b for _, b in r)
        dt = time.perf_counter() - t0
    n = per_step * steps
    return dict(value=n / dt, seconds=dt, cores=cores, problems=n, ok=oks, evals_per_s=evals / dt, per_step=per_step)


def c_port_run(wl, count=256):
    """Single-thread plain-C port (oracle/minco_oracle.c) on the first `count` problems -- a stricter CPU yardstick."""
    from oracle import c_oracle
    m = c_oracle.OracleMap.from_world(wl['world'])
    p = c_oracle.Params.from_config(wl['cfg'])
    sl = slice(0, count)
    t0 = time.perf_counter()
    out = c_oracle.plan_batch(p, m, wl['M'], wl['head'][sl], wl['tail'][sl], wl['q0'][sl], wl['ts0'][sl], wl['retry_q'][sl],
                              wl['retry_ts'], 5)
    dt = time.perf_counter() - t0
    return dict(value=count / dt, cores=1, kind='port-c', sample=f'first {count} problems, 1 thread', ok=int(out['ok'].sum()))


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_optimize launch from the committed `ncu --set full` capture
    (profiles/ncu_k_optimize.json), or None when no capture exists for this workload."""
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_k_optimize.json')))
        return d[name]['dram_bytes_read'] + d[name]['dram_bytes_write']
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap'}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=['c2', 'c4', 'c5'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    K, W = args.steps, max(args.warmup, 3 if args.impl == 'ours' else 0)

    if args.impl == 'reference':
        if rank != 0:
            return 0
        wl = cpu_sample_workload(args.workload)
        r = cpu_reference_run(args.workload, K, args.warmup)
        sample = (f"{r['per_step']} plans per step x {K} steps (problems of the same seeded workload, cycled), "
                  f"multiprocessing.Pool({r['cores']}), Python+scipy restatement of the reference (oracle/minco_ref.py)")
        line = {'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': K,
                'warmup': args.warmup, 'ms_per_step': 1e3 * r['seconds'] / K, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': describe(wl, args.gpus),
                'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': sample,
                                 'evals_per_s': r['evals_per_s'], 'ok_fraction': r['ok'] / r['problems']},
                'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return 0

    # CPU baseline first (rank 0, N = 1 only): fork the worker pool before CUDA is initialised in this process
    cpu_base = None
    if world_size == 1 and not args.no_cpu_baseline:
        wl0 = cpu_sample_workload(args.workload)
        r = cpu_reference_run(args.workload, steps=4, warmup=1)
        cpu_base = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                    'sample': f"{r['problems']} plans of the same workload (first problems, cycled), Python+scipy "
                              f"restatement of the reference (oracle/minco_ref.py), multiprocessing.Pool({r['cores']})",
                    'evals_per_s': r['evals_per_s'], 'ok_fraction': r['ok'] / r['problems'],
                    'c_port_single_thread': c_port_run(wl0)}

    import torch
    import torch.distributed as dist
    import ctypes as C
    from neo_planner_b200 import lib

    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    distributed = world_size > 1
    if distributed:
        dist.init_process_group('nccl', device_id=dev)

    wl = workload(args.workload, rank, world_size)
    cfg, M, B, world = wl['cfg'], wl['M'], wl['B'], wl['world']
    n, nq = 3 * M - 2, 2 * (M - 1)
    h = lib.Handle(cfg, local_rank, len(wl['worlds']))
    for slot, w_ in enumerate(wl['worlds']):
        h.set_map_occupancy(slot, w_.H, w_.W, w_.res, w_.ox, w_.oy, w_.occ)                # device EDT build
    ids_d = None if wl['map_ids'] is None else torch.from_numpy(wl['map_ids']).to(dev)
    fp64_peak = h.fp64_peak()

    # ---- inputs resident in HBM ------------------------------------------------------------------------------
    tau0, st0 = h.T2tau(wl['ts0'])
    rtau, rst = h.T2tau(wl['retry_ts'])
    assert not st0.any() and not rst.any()
    x0 = torch.from_numpy(np.concatenate([wl['q0'].reshape(B, nq), tau0], axis=1)).to(dev)
    head = torch.from_numpy(lib.pad_state(wl['head'])).to(dev)
    tail = torch.from_numpy(lib.pad_state(wl['tail'])).to(dev)
    rq = torch.from_numpy(wl['retry_q'].reshape(B, -1)).to(dev)
    rtau_d = torch.from_numpy(rtau).to(dev)
    # one packed result record buffer per rank: [x | ts | coeffs | costs] doubles + [status ok attempt nit runs nfev] ints
    rec_d = n + M + 12 * M + 4
    out_f = torch.zeros(B * rec_d, dtype=torch.float64, device=dev)
    out_i = torch.zeros(B * 6, dtype=torch.int32, device=dev)
    work = torch.zeros(B * 4, dtype=torch.int64, device=dev)
    off = np.cumsum([0, B * n, B * M, B * 12 * M]) * 8
    res = lib.Result()
    res.x, res.ts, res.coeffs, res.costs = (out_f.data_ptr() + int(o) for o in off)
    res.status, res.ok, res.attempt, res.nit, res.runs, res.nfev = (out_i.data_ptr() + 4 * B * i for i in range(6))
    res.work = work.data_ptr()
    if distributed:
        all_f = torch.zeros(world_size * B * rec_d, dtype=torch.float64, device=dev)
        all_i = torch.zeros(world_size * B * 6, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    stream = torch.cuda.Stream(device=dev)     # a real (non-default) stream: NULL would mean "the handle's own stream"
    torch.cuda.set_stream(stream)

    def launch_kernel():
        st = torch.cuda.current_stream().cuda_stream
        assert st != 0
        h._ck(h.lib.neo_optimize_dev(h.h, B, M, x0.data_ptr(), None, head.data_ptr(), tail.data_ptr(),
                                     None if ids_d is None else ids_d.data_ptr(), rq.data_ptr(),
                                     rtau_d.data_ptr(), 0, 5, C.byref(res), C.c_void_p(st)))

    def step_device():
        launch_kernel()
        if distributed:
            dist.all_gather_into_tensor(all_f, out_f)
            dist.all_gather_into_tensor(all_i, out_i)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        flush.zero_()
        step_device()
    barrier()
    launches0 = h.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()                      # L2 flush, outside the event bracket
        a.record()
        step_device()
        b.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = h.launch_count() - launches0        # kernels of this library launched inside the timed region
    # keep the GPU under the same load a little longer if the timed region was too short to sample clocks.
    # Kernel launches only: the number of extra iterations differs per rank, so no collective may run here.
    t_extra = time.perf_counter()
    while len(sampler.samples) < 20 and time.perf_counter() - t_extra < 2.0:
        launch_kernel()
        torch.cuda.synchronize()
    clocks = sampler.result()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    value = world_size * B * K / (total_ms * 1e-3)

    # ---- results of the last step: work accounting + sanity ----------------------------------------------------
    ok_frac = float(out_i[B:2 * B].float().mean().item())
    nfev = out_i[5 * B:6 * B].double()
    wk = work.view(B, 4).double().sum(0)
    evals = float(nfev.sum().item())
    flops = 1280.0 * M * evals + 50.0 * wk[0].item() + 65.0 * wk[1].item() + 56.0 * wk[2].item()
    l2_bytes = 8.0 * wk[0].item() + 16.0 * wk[2].item()
    hbm_bytes = B * (8 * n + 96 + 8 * (12 * M + M + 4) + 12) + len(wl['worlds']) * world.H * world.W * 32
    kern_ms = total_ms / K if not distributed else None
    # kernel-only duration (single GPU: the step IS one kernel launch + a 4-byte memset)
    step_ms = total_ms / K

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ---------------------------------
    out_host = lib.Handle.alloc_result(B, M)
    hp, tp = lib.pad_state(wl['head']), lib.pad_state(wl['tail'])
    for _ in range(2):
        h.optimize(M, wl['q0'], wl['ts0'], hp, tp, wl['map_ids'], wl['retry_q'], wl['retry_ts'], 5, out=out_host)
    barrier()
    Ke = max(3, min(K, 50))
    t0 = time.perf_counter()
    for _ in range(Ke):
        h.optimize(M, wl['q0'], wl['ts0'], hp, tp, wl['map_ids'], wl['retry_q'], wl['retry_ts'], 5, out=out_host)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world_size * B * Ke / float(t_e2e.item())
    h2d = world_size * 8 * (B * n + 12 * B + B * 4 * nq + M)                 # whole job, all ranks
    d2h = world_size * (8 * B * (n + M + 12 * M + 4) + 4 * B * 6 + 8 * B * 4)
    assert np.array_equal(out_host['ok'], out_i[B:2 * B].cpu().numpy())      # both paths computed the same thing

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        ach_tflops = flops / (step_ms * 1e-3) / 1e12
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world_size, 'steps': K, 'warmup': W,
                'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': wl['scaling'], 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic', 'config': describe(wl, world_size),
                'evals_per_s': world_size * evals / (step_ms * 1e-3), 'mean_evals_per_traj': evals / B, 'ok_fraction': ok_frac,
                'clocks': clocks,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': Ke},
                'gpu_launches': int(launches) * world_size,
                'roofline': {'bound': 'fp64', 'kernel': 'k_optimize', 'achieved': ach_tflops, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                             'frac': ach_tflops / fp64_peak, 'traffic': ncu_traffic(wl['name']),
                             'peak_source': 'measured in this run (neo_fp64_peak DFMA microbenchmark; MEASURED_PEAKS.json has no fp64 figure)',
                             'flops_per_launch': flops, 'l2_gather_bytes_per_launch': l2_bytes,
                             'hbm': {'bound': 'hbm', 'achieved': hbm_bytes / (step_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                                     'frac': hbm_bytes / (step_ms * 1e-3) / 1e9 / hbm_peak, 'algorithmic_bytes_per_launch': hbm_bytes,
                                     'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback'}},
                'wall_s_timed_region': t_wall}
        if cpu_base is not None:
            line['cpu_baseline'] = cpu_base
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())

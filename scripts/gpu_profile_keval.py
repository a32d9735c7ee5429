"""Warm-up + two k_eval launches on a bench workload (ncu target).   python scripts/gpu_profile_eval.py c4"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neo_planner_b200 import lib
import bench
name = sys.argv[1] if len(sys.argv) > 1 else 'c4'
wl = bench.workload(name, 0, 1)
dev = torch.device('cuda:0')
with torch.cuda.stream(torch.cuda.Stream()):
    run = bench.DeviceRun(wl, 0, dev, torch, lib, C)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    r, tf = bench.eval_rate(torch, run, wl, 2, 2, flush)
    print(r['ms_per_launch'], r['value'])

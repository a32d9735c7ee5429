"""A/B of two builds on one workload slice (development): python scripts/gpu_ab_two.py c4 65536 libA.so libB.so"""
import os, subprocess, sys
name, B = sys.argv[1], sys.argv[2]
for so in sys.argv[3:]:
    env = dict(os.environ, NEO_SO=os.path.abspath(so)); best = 1e9
    for rep in range(2):
        r = subprocess.run([sys.executable, 'scripts/gpu_profile_opt.py', name, B], env=env, capture_output=True, text=True, timeout=100)
        if r.returncode: print(r.stderr[-200:]); break
        best = min(best, float(r.stdout.strip().split('\n')[-1].split()[0]))
    print(name, B, os.path.basename(so), round(best, 3), flush=True)

"""A/B of one environment switch on bench workloads (development): python scripts/gpu_ab_env.py NEO_ALLPIECES=1 c2:1024 c4:4096"""
import os, subprocess, sys
switch = sys.argv[1]; k, v = switch.split('=')
for case in sys.argv[2:]:
    name, B = case.split(':')
    for on in (False, True):
        env = dict(os.environ)
        if on: env[k] = v
        best = 1e9; okf = None
        for rep in range(2):
            r = subprocess.run([sys.executable, 'scripts/gpu_profile_opt.py', name, B], env=env, capture_output=True, text=True, timeout=150)
            if r.returncode: print(r.stderr[-300:]); break
            ms, okf = r.stdout.strip().split('\n')[-1].split()
            best = min(best, float(ms))
        print(name, B, switch if on else 'default', round(best, 3), okf, flush=True)

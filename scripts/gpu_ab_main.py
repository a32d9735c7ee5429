"""A/B of two builds on the bench workloads (development): kernel ms of one neo_optimize call, best of the script's 2.
    python scripts/gpu_ab_main.py libA.so libB.so"""
import os, subprocess, sys
cases = [('c4', 65536), ('c4', 16384), ('c4', 4096), ('c2', 1024), ('c5', 16384)]
for name, B in cases:
    for so in sys.argv[1:]:
        env = dict(os.environ, NEO_SO=os.path.abspath(so))
        best = 1e9
        for rep in range(2):
            r = subprocess.run([sys.executable, 'scripts/gpu_profile_opt.py', name, str(B)], env=env, capture_output=True, text=True, timeout=150)
            if r.returncode: print(r.stderr[-300:]); break
            best = min(best, float(r.stdout.strip().split('\n')[-1].split()[0]))
        print(name, B, os.path.basename(so), round(best, 3), flush=True)

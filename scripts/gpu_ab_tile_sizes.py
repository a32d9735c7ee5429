"""Lanes per problem A/B on slices of config 4 with the shipped policy for everything else (development)."""
import os, subprocess, sys
for B in (4096, 6144, 8192, 12288, 16384):
    for tile in (8, 32):
        env = dict(os.environ, NEO_TILE=str(tile))
        best = 1e9
        for rep in range(2):
            r = subprocess.run([sys.executable, 'scripts/gpu_profile_opt.py', 'c4', str(B)], env=env, capture_output=True, text=True, timeout=150)
            if r.returncode: print(r.stderr[-300:]); break
            best = min(best, float(r.stdout.strip().split('\n')[-1].split()[0]))
        print('c4', B, 'tile', tile, round(best, 3), flush=True)

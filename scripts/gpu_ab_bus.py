"""A/B of the SM-wide (grouped departures) variant (development).  NEO_SO=... python scripts/gpu_ab_bus.py"""
import os, subprocess, sys
cases = [('c4', 65536, 0, q) for q in (0, 7, 8, 9, 10)] + [('c4', 16384, 0, q) for q in (0, 7, 8, 9)] + [('c4', 8192, t, q) for t, q in ((32, 0), (8, 8), (8, 6))] + [('c4', 4096, 8, 8)]
for name, B, tile, q in cases:
    env = dict(os.environ, NEO_GROUPED='1' if q else '0')
    if q: env['NEO_GROUP_WARPS'] = str(q)
    if tile: env['NEO_TILE'] = str(tile)
    r = subprocess.run([sys.executable, 'scripts/gpu_profile_opt.py', name, str(B)], env=env, capture_output=True, text=True, timeout=120)
    print(name, B, 'tile', tile or 'auto', 'group', q or 'off', r.stdout.strip().split('\n')[-1] if r.returncode == 0 else r.stderr[-300:], flush=True)

"""A/B of the SM-wide (grouped departures) variant (development).  NEO_SO=... python scripts/gpu_ab_bus.py"""
import os, subprocess, sys
cases = [('c5', 16384, 0, q) for q in (8, 9, 10)] + [('c5', 2048, 0, q) for q in (0, 9)] + [('c4', 2048, 0, q) for q in (0, 9)] + [('c4', 6144, 0, q) for q in (0, 9)]
for name, B, tile, q in cases:
    env = dict(os.environ)
    if q: env.update(NEO_GROUPED='1', NEO_GROUP_WARPS=str(q))
    if tile: env['NEO_TILE'] = str(tile)
    r = subprocess.run([sys.executable, 'scripts/gpu_profile_opt.py', name, str(B)], env=env, capture_output=True, text=True, timeout=120)
    print(name, B, 'tile', tile or 'auto', 'group', q or 'off', r.stdout.strip().split('\n')[-1] if r.returncode == 0 else r.stderr[-300:], flush=True)

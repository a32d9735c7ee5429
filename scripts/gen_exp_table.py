"""Prints the constants of neo_planner_b200/csrc/dd_exp.h (2^(j/32) as double-double, ln2/32 split)."""
import math
import mpmath as mp

mp.mp.prec = 400
ln2_32 = mp.log(2) / 32
e = math.floor(mp.log(ln2_32, 2))
scale = mp.mpf(2) ** (35 - e)
hi = mp.nint(ln2_32 * scale) / scale          # 36 significant bits: k*hi is exact for |k| < 2^17
print('LN2_32_HI', float(hi).hex(), 'LN2_32_LO', float(ln2_32 - hi).hex(), 'INV_LN2_32', float(32 / mp.log(2)).hex())
for j in range(32):
    v = mp.mpf(2) ** (mp.mpf(j) / 32)
    th = float(v)
    print('{%s, %s},' % (th.hex(), float(v - mp.mpf(th)).hex()))

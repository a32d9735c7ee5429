"""Instruction-footprint view of an ncu `--page source --csv --print-source sass` export: executed warp-instructions,
stall samples and no-instruction stalls per 2 KB of SASS, and the size of the code that carries 90/99 % of the work."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == 'Address')
ix = {n: i for i, n in enumerate(hdr)}
ins = []
for r in rows:
    if not r or not r[0].startswith('0x'): continue
    f = lambda k: int(r[ix[k]]) if r[ix[k]] not in ('', '-') else 0
    ins.append((int(r[0], 16), r[1].strip(), f('Instructions Executed'), f('# Samples'), f('stall_no_inst'), f('Thread Instructions Executed')))
base = ins[0][0]
tot = sum(i[2] for i in ins); ts = sum(i[3] for i in ins); tn = sum(i[4] for i in ins)
print(f'{len(ins)} instructions = {len(ins)*16/1024:.1f} KB; executed {tot/1e9:.2f} G; samples {ts}; no_inst {tn} ({100*tn/ts:.1f}%); threads/inst {sum(i[5] for i in ins)/tot:.1f}')
srt = sorted(ins, key=lambda i: -i[2]); acc = 0
for frac in (0.5, 0.8, 0.9, 0.95, 0.99):
    acc = 0
    for k, i in enumerate(srt):
        acc += i[2]
        if acc >= frac * tot: print(f'  {frac*100:.0f}% of executed instructions in {(k+1)*16/1024:.1f} KB'); break
if len(sys.argv) > 2:
    B = int(sys.argv[2])
    from collections import defaultdict
    b = defaultdict(lambda: [0, 0, 0])
    for a, s, e, sm, ni, _ in ins:
        k = (a - base) // B; b[k][0] += e; b[k][1] += sm; b[k][2] += ni
    for k in sorted(b): print(f'  +{k*B/1024:6.1f} KB exec {100*b[k][0]/tot:5.1f}%  samples {100*b[k][1]/ts:5.1f}%  no_inst {100*b[k][2]/max(1,b[k][1]):5.1f}% of its samples')

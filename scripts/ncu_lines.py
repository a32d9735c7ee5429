"""Summarise an ncu `--page source --csv --print-source cuda,sass` export: stall samples per CUDA source line."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur = None; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1]; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if cur is None or hdr is None or len(r) < 8 or r[0] == '': continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    # header has two 'Source' columns; use positional access
    idx = {name: i for i, name in enumerate(hdr)}
    s = int(r[idx['# Samples']] or 0) if r[idx['# Samples']] not in ('-', '') else 0
    ie = int(r[idx['Instructions Executed']] or 0) if r[idx['Instructions Executed']] not in ('-', '') else 0
    stalls = {}
    for name, i in idx.items():
        if name.startswith('stall_') and 'Not Issued' not in name and r[i] not in ('', '0', '-'):
            stalls[name[6:]] = int(r[i])
    out.append((cur.split('/')[-1], ln, s, ie, r[1], stalls))
tot = sum(o[2] for o in out); toti = sum(o[3] for o in out)
print('total samples', tot, 'total warp-instructions', toti)
byfile = collections.Counter()
for f, ln, s, ie, src, d in out: byfile[f] += s
print(byfile.most_common())
allst = collections.Counter()
for o in out:
    for k, v in o[5].items(): allst[k] += v
print('stall mix:', [(k, round(100*v/max(sum(allst.values()),1),1)) for k, v in allst.most_common(8)])
for f, ln, s, ie, src, d in sorted(out, key=lambda o: -o[2])[:top]:
    st = sorted(d.items(), key=lambda kv: -kv[1])[:3]
    print(f'{100*s/tot:5.1f}% {f}:{ln:4d} inst={ie:7d} {src.strip()[:100]}  {st}')

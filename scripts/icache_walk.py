"""How much code a warp walks per optimizer round: joins the ncu SASS export (executed counts, in address order) with
`nvdisasm -g -c` line markers (same instruction order) and sums, per source region, the 128-byte instruction lines
weighted by min(1, executions of the line / rounds): a line inside a loop is fetched once per round, a rare path rarely.
usage: icache_walk.py sass.csv all.sass <kernel substring>"""
import csv, re, sys, collections
sys.path.insert(0, 'scripts')
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == 'Address'); ix = {n: i for i, n in enumerate(hdr)}
ex = [(int(r[0], 16), r[1].strip(), int(r[ix['Instructions Executed']] or 0), int(r[ix['# Samples']] or 0), int(r[ix['stall_no_inst']] or 0))
      for r in rows if r and r[0].startswith('0x')]
txt = open(sys.argv[2]).read().split('\n'); key = sys.argv[3]
on = False; cur = None; src = []
for l in txt:
    if l.startswith('//--------------------- .text.'): on = key in l; continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): src.append(cur)
assert len(src) == len(ex), (len(src), len(ex))
import importlib.util
def region(f, n):
    R = {'minco_tile.cuh': [(0, 'tile/mem'), (209, 'begin/load_nodes'), (245, 'solve_nodes'), (302, 'hermite'), (340, 'sample_point'),
                            (417, 'times_from_tau'), (449, 'eval: energy'), (498, 'eval: sample loop+reduce'), (637, 'eval: adjoint'), (692, 'eval: G rows + grad_T')],
         'lbfgsb_tile.cuh': [(0, 'exact helpers'), (62, 'dcstep'), (150, 'dcsrch'), (193, 'ddot/loop_dot'), (242, 'potf2'), (295, 'lb_update'),
                             (333, 'lb_factor'), (412, 'lb_step'), (491, 'ls_load/store'), (514, 'opt_begin/gd'), (552, 'opt_advance')]}
    if f in R:
        name = R[f][0][1]
        for lo, nm in R[f]:
            if n >= lo: name = nm
        return f + ': ' + name
    return f
# rounds: executions of the most common count among instructions of times_from_tau
tt = [e[2] for e, s in zip(ex, src) if s and s[0] == 'minco_tile.cuh' and 417 <= s[1] < 449]
R = collections.Counter(tt).most_common(1)[0][0]
print('rounds (warp evaluations):', R)
base = ex[0][0]; lines = collections.defaultdict(list)
for e, s in zip(ex, src): lines[(e[0] - base) // 128].append((e, s))
walk = collections.Counter(); stat = collections.Counter(); smp = collections.Counter(); ni = collections.Counter()
for k, L in lines.items():
    w = min(1.0, max(e[2] for e, _ in L) / R)
    reg = collections.Counter(region(*s) if s else '?' for _, s in L).most_common(1)[0][0]
    walk[reg] += w * 128; stat[reg] += 128
    smp[reg] += sum(e[3] for e, _ in L); ni[reg] += sum(e[4] for e, _ in L)
tw = sum(walk.values()); ts = sum(smp.values())
print(f'walked per round: {tw/1024:.1f} KB of {sum(stat.values())/1024:.1f} KB')
for k, v in walk.most_common():
    print(f'  {k:46s} walked {v/1024:5.1f} KB  static {stat[k]/1024:5.1f} KB  samples {100*smp[k]/ts:5.1f}%  no_inst {100*ni[k]/max(1,smp[k]):3.0f}%')
# line utilisation: instructions executed at least once per two rounds / 8 per walked line
used = collections.Counter(); 
for k, L in lines.items():
    w = min(1.0, max(e[2] for e, _ in L) / R)
    if w < 0.05: continue
    reg = collections.Counter(region(*s) if s else '?' for _, s in L).most_common(1)[0][0]
    used[reg] += sum(min(1.0, e[2] / R) for e, _ in L) / (8 * w) * w * 128
print('useful bytes per round (instructions weighted by their own execution rate):', round(sum(used.values()) / 1024, 1), 'KB')
for k, v in walk.most_common(14): print(f'  {k:46s} utilisation {100*used[k]/max(1,v):4.0f}%')
if len(sys.argv) > 4:
    per = collections.Counter()
    for e, s in zip(ex, src):
        if s: per[s] += min(1.0, e[2] / R) * 16
    want = sys.argv[4]
    print('--- walked bytes per source line in', want)
    for (f, n), v in sorted(per.items(), key=lambda kv: -kv[1]):
        if (f == want or want == 'all') and v >= (160 if want == 'all' else 48): print(f'  {f}:{n:4d} {v:6.0f} B')

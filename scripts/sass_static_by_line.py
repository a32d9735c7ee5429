"""Static SASS size per source region of one kernel, from `nvdisasm -g -c` output (line markers `//## File "f", line n`).
usage: sass_static_by_line.py all.sass <mangled kernel substring> [ncu cuda-view csv]"""
import re, sys, collections, csv
txt = open(sys.argv[1]).read().split('\n'); key = sys.argv[2]
on = False; cur = None; cnt = collections.Counter(); order = []
for l in txt:
    if l.startswith('//--------------------- .text.'):
        on = key in l; continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l) and cur: cnt[cur] += 1; order.append(cur)
print('instructions', sum(cnt.values()), '=', sum(cnt.values()) * 16 / 1024, 'KB')
# regions: functions by line range (hand-maintained for the two big headers)
def region(f, n):
    R = {'minco_tile.cuh': [(0, 'tile/mem'), (209, 'begin/load_nodes'), (245, 'solve_nodes'), (302, 'hermite'), (340, 'sample_point'),
                            (417, 'times_from_tau'), (449, 'eval: energy'), (498, 'eval: sample loop+reduce'), (637, 'eval: adjoint'), (692, 'eval: G rows + grad_T')],
         'lbfgsb_tile.cuh': [(0, 'exact helpers'), (62, 'dcstep'), (150, 'dcsrch'), (193, 'ddot/loop_dot'), (242, 'potf2'), (295, 'lb_update'),
                             (333, 'lb_factor'), (412, 'lb_step'), (491, 'ls_load/store'), (514, 'opt_begin/gd'), (552, 'opt_advance')],
         }
    if f in R:
        name = R[f][0][1]
        for lo, nm in R[f]:
            if n >= lo: name = nm
        return f + ': ' + name
    return f
ex = {}
if len(sys.argv) > 3:
    rows = list(csv.reader(open(sys.argv[3]))); curf = None; hdr = None
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path': curf = r[1].split('/')[-1]; continue
        if r and r[0] == 'Line No': hdr = {n: i for i, n in enumerate(r)}; continue
        if curf and hdr and r and r[0].isdigit():
            v = r[hdr['Instructions Executed']]; s = r[hdr['# Samples']]; ni = r[hdr['stall_no_inst']]
            ex[(curf, int(r[0]))] = (int(v) if v not in ('', '-') else 0, int(s) if s not in ('', '-') else 0, int(ni) if ni not in ('', '-') else 0)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for (f, n), c in cnt.items():
    a = agg[region(f, n)]; a[0] += c
    e = ex.get((f, n), (0, 0, 0)); a[1] += e[0]; a[2] += e[1]; a[3] += e[2]
te = sum(a[1] for a in agg.values()) or 1; ts = sum(a[2] for a in agg.values()) or 1
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{k:48s} {a[0]*16/1024:6.1f} KB  exec {100*a[1]/te:5.1f}%  samples {100*a[2]/ts:5.1f}%  no_inst {100*a[3]/max(1,a[2]):4.0f}%')

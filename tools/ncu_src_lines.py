"""Stall samples of one kernel launch aggregated by the OUTERMOST source line (max line number of the inline stack) and by
SASS opcode, from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (development aid).
usage: python tools/ncu_src_lines.py src.csv file.cu [table_index] [N]"""
import csv, sys, re
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
src = open(sys.argv[2]).read().splitlines()
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N = int(sys.argv[4]) if len(sys.argv) > 4 else 45
base = sys.argv[2].split('/')[-1]
# tables: one per (file, launch); take the want-th table of this file
tabs = [i for i, r in enumerate(rows) if r and r[0] == "File Path" and r[1].endswith(base)]
t0 = tabs[want]
hdr = rows[t0 + 2]
ix = {h: i for i, h in enumerate(hdr)}
stallcols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
i = t0 + 3
cur = None
by_addr = {}
while i < len(rows) and not (rows[i] and rows[i][0] == "File Path"):
    r = rows[i]; i += 1
    if not r: continue
    if r[0] != '':
        try: cur = int(r[0])
        except ValueError: pass
        continue
    if len(r) > 2 and r[2].startswith('0x'):
        d = by_addr.setdefault(r[2], {'lines': set(), 'row': r})
        d['lines'].add(cur)
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
ops = defaultdict(lambda: [0.0, 0.0])
for a, d in by_addr.items():
    r = d['row']; ln = max(d['lines'])
    s = f(r, '# Samples'); ie = f(r, 'Instructions Executed')
    agg[ln][0] += s; agg[ln][1] += ie
    for c in stallcols: agg[ln][2][c] += f(r, c)
    op = r[3].split()[0] if not r[3].strip().startswith('@') else r[3].split()[1]
    ops[op.split('.')[0]][0] += s; ops[op.split('.')[0]][1] += ie
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print(f"total samples {tot:.0f}, warp instructions {toti:.4g}, SASS instructions {len(by_addr)}")
tots = defaultdict(float)
for v in agg.values():
    for c, x in v[2].items(): tots[c] += x
print(' '.join(f"{c[6:]}={x / tot * 100:.1f}%" for c, x in sorted(tots.items(), key=lambda kv: -kv[1])[:9]))
for ln, (s, ie, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    best = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    text = src[ln - 1].strip()[:64] if ln and ln <= len(src) else '?'
    print(f"{ln:>5} {s / tot * 100:5.2f}% inst {ie / toti * 100:5.2f}%  {text:64s} " + ' '.join(f"{k[6:]}={v / max(s, 1) * 100:.0f}%" for k, v in best))
print("-- by opcode")
for op, (s, ie) in sorted(ops.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{op:12s} samples {s / tot * 100:5.2f}%  instructions {ie / toti * 100:5.2f}%")

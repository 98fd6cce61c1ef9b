#!/usr/bin/env python
"""Executed warp-instructions per board-step by source function / by line.
usage: by_function.py src.csv dis.txt kernel-substring plies [lines-of-function]"""
import csv, re, sys
from collections import defaultdict
src_csv, dis, kern, plies = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
want_fn = sys.argv[5] if len(sys.argv) > 5 else None
files = {}
def src_lines(name):
    if name not in files:
        import glob
        hits = glob.glob('/root/repo/chinesechessai_b200/csrc/' + name)
        files[name] = open(hits[0]).read().split('\n') if hits else []
    return files[name]
def fn_of(name, line):
    best = '?'
    for i, l in enumerate(src_lines(name), 1):
        if i > line: break
        m = re.match(r'^(XQ_HD|__device__ __forceinline__|static __device__|__global__|template).*?(\w+)\(', l)
        if m and not l.startswith('template <'): best = m.group(2)
    return best
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; iE = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
sass = [(int(r[iE] or 0), int(r[iT] or 0)) for r in rows[2:] if len(r) > 5]
lines = []; cur = ('?', 0); on = False
for ln in open(dis, errors='replace'):
    if ln.startswith('.text.'): on = kern in ln; continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln): lines.append(cur)
assert len(lines) == len(sass), (len(lines), len(sass))
cache = {}
agg = defaultdict(lambda: [0, 0, 0]); byline = defaultdict(lambda: [0, 0, 0])
for l, (e, t) in zip(lines, sass):
    if l not in cache: cache[l] = fn_of(l[0], l[1]) if l[0].endswith(('.cuh', '.cu')) else l[0]
    k = cache[l]
    agg[k][0] += e; agg[k][1] += t; agg[k][2] += 1
    byline[l][0] += e; byline[l][1] += t; byline[l][2] += 1
tot = sum(a[0] for a in agg.values())
print(f"# {len(sass)} SASS instructions; {tot / plies:.0f} warp-inst per board-step")
if not want_fn:
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] / plies >= 1: print(f"{k:24s} {a[0] / plies:7.1f}/ply  lanes {a[1] / max(a[0], 1):5.1f}  sass {a[2]}")
else:
    for l, a in sorted(byline.items(), key=lambda kv: kv[0]):
        if cache[l] == want_fn and a[0] / plies >= 2:
            print(f"{l[0]}:{l[1]:4d} {a[0] / plies:6.1f}/ply lanes {a[1] / max(a[0], 1):5.1f} sass {a[2]:3d} | {src_lines(l[0])[l[1] - 1].strip()[:95]}")

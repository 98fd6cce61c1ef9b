#!/usr/bin/env python
"""Attribute an ncu SASS-level source page to CUDA source lines.

    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libxq_b200.so ; nvdisasm --print-line-info -c *.cubin > dis.txt
    python profiles/sass_by_line.py src.csv dis.txt <mangled-kernel-substring> [top_n]

The cubin travels unchanged to the GPU box, so instruction order in the ncu page equals the
disassembly order; the two are zipped by index and checked by opcode."""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis_txt, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "Thread Instructions Executed",
                                "Warp Stall Sampling (All Samples)")}
sass = [(r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]] or 0),
         int(r[ci["Thread Instructions Executed"]] or 0),
         int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)) for r in rows[2:] if len(r) > 5]

lines, cur, on = [], ("?", 0), False
for ln in open(dis_txt, errors="replace"):
    if ln.startswith(".text."):
        on = kern in ln
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(1).strip()))
assert len(lines) == len(sass), (len(lines), len(sass))
mism = sum(1 for (l, a), (b, *_r) in zip(lines, sass) if a.split()[0].lstrip("@!P0123456789 ") [:3] != b.split()[0].lstrip("@!P0123456789 ")[:3])
agg = defaultdict(lambda: [0, 0, 0, 0])
for (loc, _), (_, ie, te, st) in zip(lines, sass):
    a = agg[loc]
    a[0] += ie; a[1] += te; a[2] += st; a[3] += 1
tot = sum(a[0] for a in agg.values())
tst = sum(a[2] for a in agg.values())
print(f"# {len(sass)} SASS instructions, {tot} warp-instructions executed, opcode mismatches {mism}")
print(f"{'file:line':28s} {'warp-inst':>14s} {'%':>6s} {'lanes':>6s} {'stall%':>7s} {'sass':>5s}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{loc[0] + ':' + str(loc[1]):28s} {a[0]:14d} {100 * a[0] / tot:6.2f} {a[1] / max(a[0], 1):6.1f} "
          f"{100 * a[2] / max(tst, 1):7.2f} {a[3]:5d}")

# optional: aggregate by named line ranges  "name:lo-hi,name:lo-hi" in env XQ_RANGES (file xq_rules.cuh)
import os
if os.environ.get("XQ_RANGES"):
    rng = []
    for item in os.environ["XQ_RANGES"].split(","):
        nm, r = item.split(":")
        lo, hi = r.split("-")
        rng.append((nm, int(lo), int(hi)))
    g = defaultdict(lambda: [0, 0, 0])
    for loc, a in agg.items():
        nm = loc[0]
        if loc[0] == "xq_rules.cuh":
            nm = next((n for n, lo, hi in rng if lo <= loc[1] <= hi), "rules:other")
        g[nm][0] += a[0]; g[nm][1] += a[1]; g[nm][2] += a[2]
    print("\n# by region")
    for nm, a in sorted(g.items(), key=lambda kv: -kv[1][0]):
        print(f"{nm:28s} {a[0]:14d} {100 * a[0] / tot:6.2f}% lanes {a[1] / max(a[0], 1):5.1f} stall {100 * a[2] / max(tst, 1):6.2f}%")

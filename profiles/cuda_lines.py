#!/usr/bin/env python
"""Per-CUDA-source-line totals from an ncu report captured with --import-source on:

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X.csv
    python profiles/cuda_lines.py X.csv [top_n] [plies]

Rows that carry a line number are ncu's own aggregation of the SASS rows below them.  With
`plies` (board-steps of the profiled launch) the counts are printed per board-step."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
plies = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
cur_file, hdr = "?", None
rows = []
for r in csv.reader(open(path, errors="replace")):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in reversed(list(enumerate(r)))}
        continue
    if r[0] in ("File Name", "Function Name") or hdr is None or not r[0].strip().isdigit():
        continue
    g = lambda k: float(r[hdr[k]] or 0) if hdr.get(k) is not None and r[hdr[k]] not in ("-", "") else 0.0
    try:
        rows.append((cur_file, int(r[0]), r[1].strip(), g("Instructions Executed"),
                     g("Thread Instructions Executed"), g("Warp Stall Sampling (All Samples)")))
    except (ValueError, IndexError):
        pass  # a source line whose quotes (inline asm) defeat ncu's CSV escaping
tot = sum(x[3] for x in rows)
st = sum(x[5] for x in rows)
unit = plies if plies else 1.0
print(f"# {len(rows)} source lines with code, {tot:.0f} warp-instructions" + (f" = {tot / plies:.1f} per board-step" if plies else ""))
print(f"{'file:line':22s} {'inst' + ('/ply' if plies else ''):>12s} {'%':>6s} {'lanes':>6s} {'stall%':>7s}  source")
for f, ln, src, ie, te, ss in sorted(rows, key=lambda x: -x[3])[:top]:
    print(f"{f + ':' + str(ln):22s} {ie / unit:12.2f} {100 * ie / tot:6.2f} {te / max(ie, 1):6.1f} {100 * ss / max(st, 1):7.2f}  {src[:90]}")
byfile = defaultdict(lambda: [0.0, 0.0, 0.0])
for f, ln, src, ie, te, ss in rows:
    a = byfile[f]
    a[0] += ie; a[1] += te; a[2] += ss
print("\n# by file")
for f, a in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"{f:22s} {a[0] / unit:12.2f} {100 * a[0] / tot:6.2f} {a[1] / max(a[0], 1):6.1f} {100 * a[2] / max(st, 1):7.2f}")

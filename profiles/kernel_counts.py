#!/usr/bin/env python
"""Regenerate profiles/kernel_counts.json — the ncu-derived constants bench.py's roofline object
uses — from `ncu --page raw --csv` exports of the fused playout kernels.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > profiles/r2/X_raw.csv
    python profiles/kernel_counts.py pairs=profiles/r2/playout_sm_64k_raw.csv pair=... tpb=... warp=...

Every capture is the first timed launch of `scripts/playout_rate.py 65536 2` (seed 20260, game
ids 0..65535); the number of board-steps of that launch comes from the CPU oracle."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import xq_oracle as xo  # noqa: E402

out_path = os.path.join(ROOT, "profiles", "kernel_counts.json")
counts = json.load(open(out_path)) if os.path.exists(out_path) else {}
xo.build()
plies, _ = xo.playout_many(65536, 20260, 0, 70, 0, n_threads=os.cpu_count() or 8)
for arg in sys.argv[1:]:
    mode, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, vals = rows[0], rows[2]
    d = dict(zip(hdr, vals))
    f = lambda k: float(d[k].replace(",", ""))
    unit = dict(zip(hdr, rows[1]))
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    dram = f("dram__bytes_read.sum") * scale[unit["dram__bytes_read.sum"]] + \
        f("dram__bytes_write.sum") * scale[unit["dram__bytes_write.sum"]]
    counts[mode] = {
        "kernel": "xq::" + d["Kernel Name"].split("(")[0].replace("void ", "").replace("xq::", ""),
        "warp_inst_per_board_step": f("smsp__inst_executed.sum") / plies,
        "dram_bytes_per_launch": dram,
        "board_steps_in_launch": int(plies),
        "ncu_ms": f("gpu__time_duration.sum"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lanes_per_inst": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "registers": int(f("launch__registers_per_thread")),
        "source": os.path.relpath(path, ROOT) + " (ncu --clock-control none, 65,536 boards, seed 20260; "
                  "profiles/README.md lists the sections of each capture)"}
json.dump(counts, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(counts, indent=1))

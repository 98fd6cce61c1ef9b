"""Per-task (group, chunk) start / end / SM / warp-slot records of one queue-fed playout launch
(XQ_PLAYOUT_MODE=pairq + xq_debug_playout_timing).

usage: python scripts/playout_queue_timing.py [boards] [out.npz]   (XQ_PLAYOUT_CHUNK, XQ_PLAYOUT_CTAS_PER_SM)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["XQ_PLAYOUT_MODE"] = "pairq"
import numpy as np  # noqa: E402
import torch  # noqa: E402
from chinesechessai_b200 import _lib  # noqa: E402
from chinesechessai_b200.engine import BoardBatch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
out_path = sys.argv[2] if len(sys.argv) > 2 else None
chunk = int(os.environ.get("XQ_PLAYOUT_CHUNK", "12"))
n_chunks = 1 + (70 + chunk - 1) // chunk      # queue_max_chunks(71, chunk)
groups = (n + 15) // 16
lib = _lib.load()
dev = torch.device("cuda", 0)
bb = BoardBatch(n, device=dev, hist_cap=72)
results = torch.zeros((n, 40), dtype=torch.uint8, device=dev)
for w in range(3):
    bb.reset()
    bb.playout(900 + w, 70, results=results)
torch.cuda.synchronize()
buf = torch.zeros((groups * n_chunks, 3), dtype=torch.int64, device=dev)
bb.reset()
_lib.check(lib.xq_debug_playout_timing(buf.data_ptr()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
bb.playout(20260, 70, results=results)
b.record()
torch.cuda.synchronize()
_lib.check(lib.xq_debug_playout_timing(None))
t = buf.cpu().numpy().reshape(groups, n_chunks, 3)
valid = t[:, :, 0] > 0                                 # groups have 1 + ceil((71 - first) / chunk) chunks
t0 = t[:, :, 0][valid].min()
start, end = (t[:, :, 0] - t0) * 1e-6, (t[:, :, 1] - t0) * 1e-6
tag = t[:, :, 2]
sm, wslot, cta = tag & 0xFFFF, (tag >> 16) & 0xFFFF, tag >> 32
dur = np.where(valid, end - start, 0.0)
nxt = valid[:, 1:]
gap = (start[:, 1:] - end[:, :-1])[nxt]                # ring latency between a group's chunks
q = lambda x: [round(float(v), 3) for v in np.percentile(x, [0, 5, 25, 50, 75, 95, 100])]
total = end[valid].max()
finish = np.where(valid, end, 0).max(1)
res = {"boards": n, "chunk": chunk, "max_chunks": n_chunks, "tasks": int(valid.sum()),
       "event_ms": round(a.elapsed_time(b), 3),
       "span_ms": round(float(total), 3), "chunk_ms_pct[0,5,25,50,75,95,100]": q(dur[valid]),
       "gap_ms_pct": q(gap) if gap.size else None,
       "group_finish_ms_pct": q(finish), "group_busy_ms_pct": q(dur.sum(1)),
       "warp_busy_frac": round(float(dur.sum() / (total * min(groups, 148 * 28))), 4)}
# per warp slot (priority): mean chunk duration and number of chunks served
by_slot = {}
for s in np.unique(wslot):
    m = (wslot == s) & valid
    by_slot[int(s)] = [int(m.sum()), round(float(dur[m].mean()), 3)]
res["by_warp_slot[count, mean ms]"] = by_slot
grid = np.linspace(0, total, 41)
res["running_tasks_at_40_points"] = [int(((start <= x) & (end > x) & valid).sum()) for x in grid]
print(json.dumps(res))
if out_path:
    np.savez_compressed(out_path, start=start, end=end, sm=sm, wslot=wslot, cta=cta)

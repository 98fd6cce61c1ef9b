"""Per-warp start / end times of one fused-playout launch (xq_debug_playout_timing): where does
the launch's tail come from?  Prints the distribution of warp durations, of per-SM finish times
and the number of warps still running over time.

usage: python scripts/playout_timing.py [boards] [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from chinesechessai_b200 import _lib  # noqa: E402
from chinesechessai_b200.engine import BoardBatch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
out_path = sys.argv[2] if len(sys.argv) > 2 else None
lib = _lib.load()
dev = torch.device("cuda", 0)
bb = BoardBatch(n, device=dev, hist_cap=72)
results = torch.zeros((n, 40), dtype=torch.uint8, device=dev)
for w in range(3):
    bb.reset()
    bb.playout(900 + w, 70, results=results)
torch.cuda.synchronize()
mode = os.environ.get("XQ_PLAYOUT_MODE", "pair" if n >= 40960 else "warp")
lanes_per_board = {"pair": 2, "tpb": 1}.get(mode)
assert lanes_per_board, "timing is recorded by the per-lane kernels (XQ_PLAYOUT_MODE=pair|tpb)"
n_warps = (n * lanes_per_board + 127) // 128 * 4
buf = torch.zeros((n_warps, 3), dtype=torch.int64, device=dev)
bb.reset()
_lib.check(lib.xq_debug_playout_timing(buf.data_ptr()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
bb.playout(20260, 70, results=results)
b.record()
torch.cuda.synchronize()
_lib.check(lib.xq_debug_playout_timing(None))
t = buf.cpu().numpy().astype(np.int64)
start, end, sm = t[:, 0], t[:, 1], t[:, 2]
t0 = start.min()
start, end = (start - t0) * 1e-6, (end - t0) * 1e-6      # ms
total = end.max()
dur = end - start
q = lambda x: [round(float(v), 3) for v in np.percentile(x, [0, 5, 25, 50, 75, 95, 100])]
sm_end = np.array([end[sm == s].max() for s in np.unique(sm)])
sm_cnt = np.array([(sm == s).sum() for s in np.unique(sm)])
grid = np.linspace(0, total, 41)
running = [int(((start <= x) & (end > x)).sum()) for x in grid]
res = {"mode": mode, "boards": n, "warps": int(n_warps), "event_ms": round(a.elapsed_time(b), 3),
       "span_ms": round(float(total), 3),
       "warp_start_ms_pct[0,5,25,50,75,95,100]": q(start), "warp_end_ms_pct": q(end),
       "warp_duration_ms_pct": q(dur), "sm_finish_ms_pct": q(sm_end),
       "sm_idle_frac": round(float(1 - sm_end.mean() / total), 4),
       "warps_per_sm_min_max": [int(sm_cnt.min()), int(sm_cnt.max())], "sms": int(len(sm_cnt)),
       "mean_running_warps_frac": round(float(dur.sum() / (total * n_warps)), 4),
       "running_warps_at_40_points": running}
print(json.dumps(res))
if out_path:
    np.savez_compressed(out_path, start=start, end=end, sm=sm)

"""Time the fused playout kernel alone (cfg 2 shape by default) and print the rate together with
an XOR of the per-game digests, so that two builds can be compared for speed AND identity.

usage: python scripts/playout_rate.py [boards] [launches]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from chinesechessai_b200.engine import BoardBatch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
bb = BoardBatch(n, device=dev, hist_cap=72)
results = torch.zeros((n, 40), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for w in range(3):
    bb.reset()
    bb.playout(900 + w, 70, results=results)
torch.cuda.synchronize()
ms, plies, x = 0.0, 0, 0
for k in range(reps):
    flush.zero_()
    bb.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    bb.playout(20260 + k, 70, results=results)
    b.record()
    torch.cuda.synchronize()
    ms += a.elapsed_time(b)
    plies += int(results.view(torch.int32)[:, 0].sum())
    d = results.view(torch.int64)[:, 3].cpu().numpy().view("u8")
    acc = 0
    for v in d.tolist():
        acc ^= v
    x ^= acc
print(json.dumps({"mode": os.environ.get("XQ_PLAYOUT_MODE", "default"), "boards": n,
                  "ms_per_launch": round(ms / reps, 3),
                  "board_steps_per_s": round(plies / (ms * 1e-3)), "digest_xor": hex(x)}))

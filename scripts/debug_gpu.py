import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chinesechessai_b200 import engine as eng
from oracle import xq_oracle as xo
bb = eng.BoardBatch(1)
e = xo.Env()
for mvt in [(9,1,7,2),(0,1,2,2),(7,2,9,1)]:
    mv = torch.tensor([eng.pack_move(mvt)], dtype=torch.int16, device=bb.device)
    r, f = bb.step(mv)
    ro = e.make_move(mvt)
    m = bb.meta_host()[0]
    print("gpu", float(r[0]), int(f[0]), m, "| oracle", ro, e.s.consecutive_checks, e.s.no_capture)
    print(" hist gpu", bb.pos_hist_host()[0,:3], "oracle", e.position_history[:3])
res, tr = eng.BoardBatch(2).playout(0x5EED, 6, trace=True)
print(eng.results_host(res))
print(tr["reward"].cpu().numpy(), tr["flags"].cpu().numpy(), tr["pick"].cpu().numpy(), tr["n"].cpu().numpy())
e = xo.Env(); r, t = e.playout(0x5EED, 0, 6, 0, trace=True)
print(r.plies, r.digest, t["reward"], t["flags"], t["pick"], t["n"])

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chinesechessai_b200 import engine as eng
SEED = 0xC0FFEE
for n, bias in ((65536, 0), (50000, 160)):
    fused = eng.BoardBatch(n)
    rf = eng.results_host(fused.playout(SEED, 70, capture_bias=bias))
    bb = eng.BoardBatch(n)
    plies = torch.zeros(n, dtype=torch.int32, device=bb.device)
    rsum = torch.zeros(n, dtype=torch.float64, device=bb.device)
    bb.legal_moves()
    for ply in range(70):
        mv = bb.pick(SEED, ply, capture_bias=bias)
        plies += (mv >= 0).to(torch.int32)
        reward, flags = bb.step(mv, want_next=True)
        rsum += torch.where(mv >= 0, reward, torch.zeros_like(reward))
    ok = (np.array_equal(plies.cpu().numpy(), rf["plies"]) and
          np.array_equal(rsum.cpu().numpy().view(np.uint64), rf["reward_sum"].view(np.uint64)) and
          np.array_equal(bb.boards_host(), fused.boards_host()) and
          np.array_equal(bb.meta_host(), fused.meta_host()) and
          np.array_equal(bb.pos_hist_host()[:, :70], fused.pos_hist_host()[:, :70]))
    print(n, bias, "step-per-launch (pair-mapped) == fused playout:", ok, int(rf["plies"].sum()), "plies")

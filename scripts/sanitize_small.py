"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chinesechessai_b200.engine import BoardBatch, encode_planes, policy_priors, results_host, playout_host
from chinesechessai_b200.mcts import BatchedMCTS, HashEvaluator
n = 64
bb = BoardBatch(n)
bb.legal_moves()
for ply in range(6):
    mv = bb.pick(3, ply, capture_bias=128)
    bb.step(mv, want_next=True)
r, tr = bb.playout(5, 70, capture_bias=100, trace=True)
print("plies", int(results_host(r)["plies"].sum()))
bb2 = BoardBatch(n); bb2.playout(9, 20)
m = BatchedMCTS(n, 30)
mvs, vis, nc = m.search(bb2.board, bb2.meta, HashEvaluator())
print("visits", int(vis.sum()))
pl = bb2.meta[:, 0].view(torch.int8)
x = encode_planes(bb2.board, pl)
mv, nm = bb2.legal_moves()
p = policy_priors(torch.randn(n, 8100, device="cuda"), mv, nm)
q = torch.zeros((n, 4), dtype=torch.uint8, device="cuda")
bb2.lib.xq_query_checks(bb2.board.data_ptr(), bb2.meta.data_ptr(), q.data_ptr(), n, None)
print("hash", int(bb2.position_hash()[0]))
init = BoardBatch(n)
b0 = np.ascontiguousarray(init.board.cpu().numpy()); m0 = np.ascontiguousarray(init.meta_host())
print("host", int(playout_host(b0, m0, 1, 30)["plies"].sum()))
torch.cuda.synchronize()
print("ok")

"""A few batched MCTS self-play plies (cfg 3 shape) — target for the ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chinesechessai_b200.neural_network import ChessNet
from chinesechessai_b200.self_play import BatchedSelfPlay
torch.manual_seed(0)
net = ChessNet().cuda().eval()
sp = BatchedSelfPlay(net, 4096, 15, temperature=1.0, net_dtype=torch.bfloat16, seed=0)
sp.boards.playout(0x5EED, 4)
sp.play(int(sys.argv[1]) if len(sys.argv) > 1 else 3, check_done=False)
torch.cuda.synchronize()
print("plies", sp.plies, sp.stats())

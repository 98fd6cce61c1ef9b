"""Latency of the single-game call surface (what evaluate.py / compare_models.py drive):
MCTS(net, n).search(env) per move and self_play_game() per game, next to one batched ply at
batch 1 and the reference's own CPU numbers (32 s per 15-sim game)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chinesechessai_b200.chess_env import ChineseChess
from chinesechessai_b200.neural_network import ChessNet
from chinesechessai_b200.self_play import MCTS, self_play_game, BatchedSelfPlay

sims = int(sys.argv[1]) if len(sys.argv) > 1 else 15
torch.manual_seed(0)
net = ChessNet().cuda().eval()
env = ChineseChess()
m = MCTS(net, num_simulations=sims)
m.search(env); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    m.search(env)
torch.cuda.synchronize()
t_search = (time.perf_counter() - t0) / 20
np.random.seed(0)
self_play_game(net, temperature=1.0, num_simulations=sims)
t0 = time.perf_counter()
data, winner, reason = self_play_game(net, temperature=1.0, num_simulations=sims)
t_game = time.perf_counter() - t0
out = {"sims": sims, "search_ms_per_move": 1e3 * t_search, "self_play_game_s": t_game, "plies": len(data)}
for n in (1, 64, 512, 1024, 2048):
    for graph in (False, True):
        sp = BatchedSelfPlay(net, n, sims, 1.0, net_dtype=torch.bfloat16, seed=0, use_graph=graph)
        sp.play(3, check_done=False); torch.cuda.synchronize()
        t0 = time.perf_counter()
        sp.play(20, check_done=False); torch.cuda.synchronize()
        out[f"batched_ply_ms_n{n}_{'graph' if graph else 'eager'}"] = round(1e3 * (time.perf_counter() - t0) / 20, 3)
print(json.dumps(out))

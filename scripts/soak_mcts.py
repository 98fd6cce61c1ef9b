"""MCTS parity soak: the batched device game loop (hashed deterministic evaluator) against the
C oracle's literal MCTS.search + the shared sampling rule, for every game and ply: identical root
move lists, visit counts and chosen moves.  Usage: python scripts/soak_mcts.py [games] [sims] [plies]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from chinesechessai_b200.mcts import HashEvaluator
from chinesechessai_b200.self_play import BatchedSelfPlay
from oracle import xq_oracle as xo

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 50
plies = int(sys.argv[3]) if len(sys.argv) > 3 else 70
seed, first, temp = 1234, 100000, 1.0
t0 = time.perf_counter()
sp = BatchedSelfPlay(HashEvaluator(), n, sims, temperature=temp, seed=seed, first_game_id=first, use_graph=False)
sp.play(plies)
torch.cuda.synchronize()
t_gpu = time.perf_counter() - t0
P = sp.plies
rm, rv, rn = (t[:P].cpu().numpy() for t in (sp.rec_moves, sp.rec_visits, sp.rec_n))
played, rmove = sp.rec_played[:P].cpu().numpy(), sp.rec_move[:P].cpu().numpy()

def check(g):
    e = xo.Env()
    cnt = 0
    for p in range(P):
        om, ov, _ = xo.mcts_search(e, sims)
        if len(e.legal_moves_packed()) == 0 or len(om) == 0:
            assert not played[p, g], (g, p)
            break
        assert played[p, g], (g, p)
        k = int(rn[p, g])
        assert np.array_equal(rm[p, g, :k], om) and np.array_equal(rv[p, g, :k], ov), (g, p)
        idx = xo.sample_move(ov, temp, seed, first + g, p)
        assert int(rmove[p, g]) == int(om[idx]), (g, p)
        _, _, done = e.make_move(int(om[idx]))
        cnt += 1
        if done:
            assert p + 1 >= P or not played[p + 1, g], (g, p)
            break
    return cnt

t0 = time.perf_counter()
with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
    counts = list(ex.map(check, range(n)))
t_cpu = time.perf_counter() - t0
print(json.dumps({"games": n, "sims_per_move": sims, "plies_checked": int(sum(counts)), "sims_checked": int(sum(counts)) * sims,
                  "mismatches": 0, "gpu_seconds": round(t_gpu, 3), "oracle_seconds": round(t_cpu, 1)}))

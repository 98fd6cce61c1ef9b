"""Large-volume parity soak: fused CUDA playouts vs the C oracle, every game's digest chain
(ordered move lists, picks, boards, float64 rewards, flags per ply), final hash, plies, winner,
reason and reward sum.  Usage: python scripts/soak_parity.py [games_per_chunk] [chunks]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chinesechessai_b200.engine import BoardBatch, results_host
from oracle import xq_oracle as xo

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
threads = os.cpu_count() or 1
bb = BoardBatch(n, hist_cap=72)
tot_games = tot_plies = 0
hist = np.zeros(9, np.int64)
t_gpu = t_cpu = 0.0
for c in range(chunks):
    seed = 0xABCDEF + 7919 * c
    bias = [0, 0, 64, 128, 192, 240, 0, 255][c % 8]
    first = c * n
    bb.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = results_host(bb.playout(seed, 70, first_game_id=first, capture_bias=bias))
    t_gpu += time.perf_counter() - t0
    t0 = time.perf_counter()
    total, ref = xo.playout_many(n, seed, first, 70, bias, n_threads=threads)
    t_cpu += time.perf_counter() - t0
    for f in ("plies", "winner", "reason", "max_legal", "digest", "final_hash"):
        bad = np.nonzero(res[f] != ref[f])[0]
        assert len(bad) == 0, (c, f, bad[:5])
    assert np.array_equal(res["reward_sum"].view(np.uint64), ref["reward_sum"].view(np.uint64)), c
    assert int(bb.meta_host()["flags"].max()) == 0
    tot_games += n; tot_plies += total
    hist += np.bincount(ref["reason"], minlength=9)
    print(f"chunk {c}: bias {bias:3d} {total} plies ok", flush=True)
print(json.dumps({"games": tot_games, "plies": int(tot_plies), "mismatches": 0,
                  "reason_histogram": {str(k): int(v) for k, v in enumerate(hist)},
                  "gpu_seconds_incl_copies": round(t_gpu, 3), "oracle_seconds": round(t_cpu, 1),
                  "oracle_threads": threads}))

#!/usr/bin/env python
"""Drives the reference's UNCHANGED consumer modules (trainer.py, evaluate.py, compare_models.py)
in the current Python path set-up and prints one JSON line.

Run it two ways (see baseline/reference.py):
  * env_for_shims():      chess_env / self_play / neural_network resolve to the drop-in shims
                          (integration/shims -> chinesechessai_b200), everything else to the
                          reference checkout  -> the GPU engine under the stock trainer;
  * env_for_reference():  everything resolves to the stock reference with CUDA hidden -> the
                          CPU baseline of cfg 5.
cwd should be a scratch directory: the trainer writes logs/, models/, data/ relative to it.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import time


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["consumers", "cfg5", "readers"], default="consumers")
    ap.add_argument("--games", type=int, default=8)
    ap.add_argument("--sims", type=int, default=15)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--eval-games", type=int, default=2)
    ap.add_argument("--match-games", type=int, default=2)
    ap.add_argument("--warm", type=int, default=0, help="untimed warm-up games before cfg5's timed iteration")
    a = ap.parse_args()

    import numpy as np
    import torch
    out = {"mode": a.mode}
    chatter = io.StringIO()
    if a.mode == "readers":
        return readers(out, chatter)
    with contextlib.redirect_stdout(chatter):
        import config
        import chess_env
        import neural_network
        import self_play
        import trainer
        out["modules"] = {m.__name__: os.path.dirname(os.path.abspath(m.__file__))
                          for m in (config, chess_env, neural_network, self_play, trainer)}
        out["engine"] = chess_env.ChineseChess.__module__
        out["device"] = config.DEVICE
        np.random.seed(0)
        torch.manual_seed(0)
        t = trainer.Trainer()

        if a.mode == "consumers":
            import compare_models
            import evaluate
            stats = t.collect_self_play_data(a.games)          # trainer.py:147-296 -> parallel_self_play
            out["collect"] = {k: (float(v) if isinstance(v, float) else int(v)) for k, v in stats.items()}
            out["buffer"] = len(t.replay_buffer)
            out["total_games"] = int(t.total_games)
            b0 = t.replay_buffer.buffer[0]
            out["sample"] = {"board_shape": list(np.asarray(b0[0]).shape), "board_dtype": str(np.asarray(b0[0]).dtype),
                             "n_probs": len(b0[1]), "probs_sum": float(sum(b0[1].values())),
                             "reward_type": type(b0[2]).__name__}
            out["train_loss"] = float(t.train_network())       # trainer.py:298-362
            out["training_steps"] = int(t.training_steps)
            t._log_progress(1, stats)                           # trainer.py:395-431
            t.save_model()                                      # trainer.py:433-449
            out["log_line"] = open(os.path.join(config.LOG_DIR, "training.log"), encoding="utf-8").read().strip()
            res = evaluate.evaluate_model(config.LATEST_MODEL, num_games=a.eval_games, verbose=False)  # evaluate.py:13-132
            out["evaluate"] = {k: res[k] for k in ("red_wins", "black_wins", "draws", "avg_moves", "min_moves", "max_moves")}
            net2 = neural_network.ChessNet().to(config.DEVICE)
            net2.eval()
            t.network.eval()
            out["play_match"] = compare_models.play_match(t.network, net2, num_games=a.match_games, verbose=False)
            out["best_games_pkl"] = os.path.exists(os.path.join(config.DATA_DIR, "best_games.pkl"))
        else:
            # cfg 5: one iteration = parallel_self_play (self_play.py:368) + replay push + train_network
            t.network.eval()
            if a.warm:
                self_play.parallel_self_play(t.network, num_games=a.warm, temperature=1.0,
                                             num_simulations=a.sims, num_workers=a.workers)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            results = self_play.parallel_self_play(t.network, num_games=a.games, temperature=1.0,
                                                   num_simulations=a.sims, num_workers=a.workers)
            t1 = time.perf_counter()
            plies = 0
            for game_data, winner, end_reason in results:
                t.replay_buffer.push(game_data)
                t.total_games += 1
                plies += len(game_data)
            t2 = time.perf_counter()
            loss = float(t.train_network()) if len(t.replay_buffer) >= config.BATCH_SIZE else None
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            t3 = time.perf_counter()
            out.update({"games": len(results), "sims": a.sims, "workers": a.workers, "plies": plies,
                        "self_play_s": t1 - t0, "push_s": t2 - t1, "train_s": t3 - t2, "seconds": t3 - t0,
                        "train_batches": int(t.training_steps), "loss": loss,
                        "decisive": sum(1 for _, w, _ in results if w != 0),
                        "cores": os.cpu_count(), "torch_threads": torch.get_num_threads()})
        t.close()
    sys.stderr.write(chatter.getvalue()[-4000:])
    print(json.dumps(out, ensure_ascii=False))


def readers(out, chatter) -> None:
    """The reference's own READERS on files written by chinesechessai_b200.train_loop in cwd:
    Trainer.__init__ -> load_model (trainer.py:77-78,451-459), plot_progress.parse_training_log
    (:16-63), view_best_games.load_best_games / list_best_games (:15-80) and the move replay of
    GameReplayer.replay_game (:201-213).  matplotlib and pygame are not in this image; the
    functions used here never touch them, so empty stand-in modules satisfy the imports."""
    import types
    import numpy as np
    for name in ("matplotlib", "matplotlib.pyplot", "pygame"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "rcParams"):
        sys.modules["matplotlib"].rcParams = {}
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    with contextlib.redirect_stdout(chatter):
        import config
        import chess_env
        import plot_progress
        import trainer
        import view_best_games
        out["modules"] = {m.__name__: os.path.dirname(os.path.abspath(m.__file__))
                          for m in (config, chess_env, trainer, plot_progress, view_best_games)}
        t = trainer.Trainer()                                   # loads models/latest.pt
        out["total_games"], out["training_steps"] = int(t.total_games), int(t.training_steps)
        out["adam_steps"] = sorted({int(v["step"]) for v in t.optimizer.state_dict()["state"].values()})
        t.close()
        data = plot_progress.parse_training_log(os.path.join(config.LOG_DIR, "training.log"))
        out["log"] = data
        games = view_best_games.load_best_games()
        view_best_games.list_best_games(games)
        out["best_games"] = len(games)
        replayed = []
        for g in games[:4]:                                     # view_best_games.py:201-213
            env = chess_env.ChineseChess()
            env.reset()
            n = 0
            for board_state, move_probs, player in g["game_data"]:
                assert np.array_equal(np.asarray(board_state), env.board), "recorded board != replayed board"
                if move_probs:
                    # the viewer replays the arg-max move; the game itself sampled, so follow
                    # the recorded boards instead and only check that the arg-max is legal
                    best_move = max(move_probs.items(), key=lambda x: x[1])[0]
                    assert best_move in env.get_legal_moves()
                    assert type(move_probs) is dict and abs(sum(move_probs.values()) - 1.0) < 1e-9
                n += 1
                break                                           # boards after ply 0 depend on the sampled move
            replayed.append({"winner": int(g["winner"]), "moves": int(g["moves"]), "type": g["type"],
                             "samples": len(g["game_data"]), "timestamp": g["timestamp"].isoformat()})
        out["replayed"] = replayed
    sys.stderr.write(chatter.getvalue()[-3000:])
    print(json.dumps(out, ensure_ascii=False))


if __name__ == "__main__":
    main()

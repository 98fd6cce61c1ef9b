"""Batched ``play_match`` (reference: compare_models.py:13-92): model 1 plays red, model 2 black,
``num_games`` games, move choice by temperature 0.3 over the visit counts.

The reference plays the games one after the other with one ``MCTS.search`` (one forward per
wave) per move; here all games of the match run as ONE device batch — red's positions go through
``network1``, black's through ``network2``, the same kernels as ``parallel_self_play`` in
opponent mode.  Same result dict.  Differences a caller can observe: moves are drawn from the
engine's counter-based stream (seeded once from ``np.random``) instead of one ``np.random.choice``
per move, and the per-game progress lines are printed after the batch instead of between games.
The unmodified ``compare_models.play_match`` itself also runs on the drop-in ``ChineseChess`` /
``MCTS`` classes (tests/test_reference_consumers.py); this module is the fast path for large
matches.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .config import MCTS_SIMULATIONS
from .self_play import BatchedSelfPlay

MATCH_TEMPERATURE = 0.3  # compare_models.py:54


def play_match(network1, network2, num_games: int = 20, verbose: bool = True,
               num_simulations: Optional[int] = None, seed: Optional[int] = None) -> Dict[str, float]:
    n_sims = num_simulations if num_simulations else MCTS_SIMULATIONS  # MCTS(network) default, self_play.py:86
    sp = BatchedSelfPlay(network1, num_games, n_sims, temperature=MATCH_TEMPERATURE,
                         opponent_network=network2, seed=seed)
    sp.play()  # the reference allows 100 plies per game; the 70-ply cap of make_move ends them first
    torch.cuda.synchronize(sp.device)
    meta = sp.boards.meta_host()
    w = meta["winner"].astype(np.int64)
    winner = np.where(w == _lib.WINNER_NONE, 0, w)
    moves = meta["move_count"].astype(np.int64)
    model1_wins, model2_wins = int((winner == 1).sum()), int((winner == -1).sum())
    draws = int(num_games - model1_wins - model2_wins)
    if verbose:
        for g in range(num_games):
            result = "模型1胜" if winner[g] == 1 else "模型2胜" if winner[g] == -1 else "和局"
            print(f"  对局 {g + 1}/{num_games}... {result} ({int(moves[g])}步)")
    return {
        "model1_wins": model1_wins, "model2_wins": model2_wins, "draws": draws,
        "avg_moves": float(moves.sum()) / num_games,
        "model1_winrate": model1_wins / num_games * 100,
        "model2_winrate": model2_wins / num_games * 100,
        "draw_rate": draws / num_games * 100,
    }

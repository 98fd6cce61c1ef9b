"""Device-resident training samples.

``BatchedSelfPlay`` records boards, players, visit counts and step rewards in HBM.  This module
turns them into the tensors ``Trainer.train_network`` consumes (trainer.py:311-321: planes of
``encode_board(board, 1)`` — the player flag is hard-wired to 1 there — and the shaped scalar
reward of self_play.py:262-310) without a host round trip, and gathers them across ranks."""
from __future__ import annotations

from typing import Dict

import torch

from .engine import encode_planes
from .self_play import BatchedSelfPlay


def _final_reward(winner: torch.Tensor, player: torch.Tensor, length: torch.Tensor) -> torch.Tensor:
    """self_play.py:268-298, vectorised in float64 (same constants, same branches).  Every constant
    is a float64 tensor: ``torch.where(cond, 0.3, 0.1)`` on Python scalars would come out as
    float32 and 1.0 + 0.3 would no longer equal the reference's double."""
    dev = winner.device

    def c(x: float) -> torch.Tensor:
        return torch.tensor(x, dtype=torch.float64, device=dev)
    long_game = length >= 60
    draw = torch.where(long_game, torch.where(player == 1, c(-0.15), c(0.05)),
                       torch.where(player == 1, c(-0.1), c(0.1)))
    bonus = torch.where(length <= 30, c(0.5), torch.where(length <= 50, c(0.3),
                        torch.where(length <= 70, c(0.1), c(0.0))))
    win = c(1.0) + bonus
    lose = torch.where(long_game, c(-1.2), c(-1.0))
    return torch.where(winner == 0, draw, torch.where(winner == player, win, lose))


def training_tensors(sp: BatchedSelfPlay, red_only: bool = False) -> Dict[str, torch.Tensor]:
    """All samples of a finished ``BatchedSelfPlay`` as device tensors, game-major / ply-minor
    (the order ``materialise()`` produces): ``board`` int8[N,90], ``player`` int8[N],
    ``reward`` float64[N] (== the third field of the reference's sample tuples), ``game`` int64[N],
    ``ply`` int64[N]."""
    P, n = sp.plies, sp.n
    played = sp.rec_played[:P]                                   # [P, n]
    player = sp.rec_player[:P]
    keep = played & ((player == 1) if red_only else torch.ones_like(played))
    meta = sp.boards.meta
    w = meta[:, 1].view(torch.int8).to(torch.int64)
    winner = torch.where(w == 2, torch.zeros_like(w), w)          # None -> 0 (self_play.py:259)
    length = keep.sum(0).to(torch.int64)                          # samples per game (:264)
    # step_rewards is indexed by SAMPLE index (:303-304): the i-th kept sample of a game gets the
    # reward of the game's i-th ply
    sample_idx = keep.to(torch.int64).cumsum(0) - 1               # [P, n]
    step = sp.rec_reward[:P]                                      # reward of ply p
    imm = torch.gather(step, 0, sample_idx.clamp(min=0))
    imm = torch.where(sample_idx < played.sum(0, keepdim=True), imm, torch.zeros_like(imm))
    fin = _final_reward(winner[None, :].expand(P, n), player.to(torch.int64), length[None, :].expand(P, n))
    total = fin + imm * 0.01
    order = keep.t().reshape(-1).nonzero(as_tuple=False).squeeze(1)   # game-major
    g, p = order // P, order % P
    return {"board": sp.rec_board[:P].permute(1, 0, 2).reshape(n * P, 90)[order].contiguous(),
            "player": player.t().reshape(-1)[order].contiguous(),
            "reward": total.t().reshape(-1)[order].contiguous(), "game": g, "ply": p}


def training_batch(samples: Dict[str, torch.Tensor], index: torch.Tensor,
                   dtype: torch.dtype = torch.float32):
    """(states [B,15,10,9], target_values [B,1]) for the sample rows ``index``; the planes come
    from the CUDA encode kernel with the player flag fixed to 1 as in trainer.py:317."""
    boards = samples["board"][index].contiguous()
    ones = torch.ones((boards.shape[0],), dtype=torch.int8, device=boards.device)
    return encode_planes(boards, ones, dtype=dtype), samples["reward"][index].to(torch.float32).unsqueeze(1)

"""Training loop for one process per GPU, writing the files the reference's tools read.

The reference's loop is ``Trainer.train_loop`` (trainer.py:364-416): collect self-play games,
update the network, append a line to ``logs/training.log``, save ``models/latest.pt``, keep the
decisive games in ``data/best_games.pkl``.  That loop runs unchanged on the drop-in classes
(tests/test_reference_consumers.py).  This module is the multi-GPU form of the same loop built on
``distributed_self_play`` / ``train_on_samples``: games are sharded over the ranks, samples stay
on the device, and the trainer rank writes the three files in the reference's formats
(``formats.py``) so that ``plot_progress.py``, ``view_best_games.py``, ``evaluate.py`` and
``Trainer.load_model`` keep working on what it produces.

Launch: ``python -m chinesechessai_b200.train_loop --iterations 3 --games 4096`` on one GPU, or
under ``python -m torch.distributed.run --nproc-per-node N -m chinesechessai_b200.train_loop ...``.
"""
from __future__ import annotations

import os
import pickle
import time
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, formats
from .config import MCTS_SIMULATIONS
from .dist import distributed_self_play
from .iteration import train_on_samples
from .mcts import NetEvaluator

LEARNING_RATE = 0.001     # config.py:35
WEIGHT_DECAY = 1e-4       # config.py:36
MAX_BEST_GAMES = 500      # trainer.py:498
BEST_PER_ITERATION = 64   # a batch of thousands of games would otherwise flood the 500-game file
SHORT_DRAW = 50           # trainer.py:239: draws shorter than this are kept as well


def _evaluator(network, precision: str):
    if precision == "bf16":
        return NetEvaluator(network, torch.bfloat16)
    if precision == "tf32":
        return NetEvaluator(network, torch.float32, tf32=True)
    if precision == "fp32":
        return NetEvaluator(network, torch.float32)
    raise ValueError(f"precision {precision!r}: expected fp32, tf32 or bf16")


class TrainLoop:
    """State of a run: network, optimizer, counters, output directories (``models/``, ``logs/``,
    ``data/`` under ``out_dir`` — the reference's MODEL_DIR / LOG_DIR / DATA_DIR, config.py:46-52)."""

    def __init__(self, network: torch.nn.Module, out_dir: str, precision: str = "fp32",
                 num_simulations: int = MCTS_SIMULATIONS, temperature: float = 1.0,
                 optimizer: Optional[torch.optim.Optimizer] = None, trainer_rank: int = 0, seed: int = 0):
        self.network = network
        self.optimizer = optimizer or torch.optim.Adam(network.parameters(), lr=LEARNING_RATE,
                                                       weight_decay=WEIGHT_DECAY)   # trainer.py:95-99
        self.evaluator = _evaluator(network, precision)
        self.precision = precision
        self.num_simulations = int(num_simulations)
        self.temperature = float(temperature)
        self.out_dir = out_dir
        self.trainer_rank = int(trainer_rank)
        self.seed = int(seed)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.total_games = 0
        self.training_steps = 0
        self.iteration = 0
        self.short_draw = SHORT_DRAW

    # -- files ------------------------------------------------------------------------------
    @property
    def latest_path(self) -> str:
        return os.path.join(self.out_dir, "models", "latest.pt")

    @property
    def log_path(self) -> str:
        return os.path.join(self.out_dir, "logs", "training.log")

    @property
    def best_games_path(self) -> str:
        return os.path.join(self.out_dir, "data", "best_games.pkl")

    def save(self) -> None:
        """``Trainer.save_model`` (trainer.py:433-449) on the trainer rank."""
        if self.rank != self.trainer_rank:
            return
        os.makedirs(os.path.dirname(self.latest_path), exist_ok=True)
        ck = formats.checkpoint_dict(self.network, self.optimizer, self.total_games, self.training_steps)
        tmp = self.latest_path + ".tmp"
        torch.save(ck, tmp)
        os.replace(tmp, self.latest_path)           # a killed run never leaves half a checkpoint
        if self.total_games % 1000 == 0:            # trainer.py:446-449
            torch.save(ck, os.path.join(self.out_dir, "models", f"model_{self.total_games}.pt"))

    def resume(self) -> bool:
        """``Trainer.load_model`` (trainer.py:451-459) on every rank (the counters must agree; the
        weights are broadcast again at the start of each iteration anyway)."""
        if not os.path.exists(self.latest_path):
            return False
        dev = next(self.network.parameters()).device
        ck = torch.load(self.latest_path, map_location=dev, weights_only=False)
        self.total_games, self.training_steps = formats.load_checkpoint(ck, self.network, self.optimizer)
        return True

    def _append_log(self, stats: Dict[str, float], samples: int) -> None:
        os.makedirs(os.path.dirname(self.log_path), exist_ok=True)
        with open(self.log_path, "a", encoding="utf-8") as f:
            f.write(formats.training_log_line(self.iteration, self.total_games, stats, samples))

    def _save_best_games(self, sp) -> int:
        """``Trainer._save_best_games`` (trainer.py:468-502) for this rank's decisive games and
        short draws (:237-240)."""
        meta = sp.boards.meta_host()
        winner = np.where(meta["winner"] == _lib.WINNER_NONE, 0, meta["winner"]).astype(np.int64)
        pick = np.flatnonzero((winner != 0) | (meta["move_count"] < self.short_draw))[:BEST_PER_ITERATION]
        if len(pick) == 0:
            return 0
        games = sp.materialise(games=pick.tolist())
        records: List[dict] = []
        if os.path.exists(self.best_games_path):
            try:
                with open(self.best_games_path, "rb") as f:
                    records = pickle.load(f)
            except Exception:
                records = []
        for (game_data, w, reason) in games:
            records.append(formats.best_game_record(game_data, w, len(game_data), reason, self.total_games))
        records = records[-MAX_BEST_GAMES:]
        os.makedirs(os.path.dirname(self.best_games_path), exist_ok=True)
        with open(self.best_games_path, "wb") as f:
            pickle.dump(records, f)
        return len(games)

    # -- one iteration ------------------------------------------------------------------------
    def step(self, num_games: int) -> Dict[str, float]:
        """All ranks call this: one iteration of trainer.py:364-416."""
        dev = next(self.network.parameters()).device
        self.iteration += 1
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        samples, sp = distributed_self_play(self.evaluator, num_games, self.num_simulations,
                                            self.temperature, seed=self.seed + self.iteration,
                                            network=self.network, dst=self.trainer_rank)
        counts = sp.outcome_counts()
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(counts)                  # 40 bytes: the iteration's statistics
        red, black, draws, plies, games = (int(x) for x in counts.tolist())
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        self.total_games += games
        stats = {"red_wins": red, "black_wins": black, "draws": draws,
                 "avg_moves": plies / games if games else 0.0}
        out: Dict[str, float] = dict(stats, iteration=self.iteration, games=games, plies=plies,
                                     self_play_s=t1 - t0, loss=None, best_games=0)
        if self.rank == self.trainer_rank:
            n_samples = int(samples["reward"].shape[0])
            loss = train_on_samples(self.network, self.optimizer, samples)
            torch.cuda.synchronize(dev)
            self.network.eval()
            out["loss"] = loss
            out["samples"] = n_samples
            self.training_steps += min(50, n_samples // 64)
            out["best_games"] = self._save_best_games(sp)
            self._append_log(stats, n_samples)
        self.save()
        out["train_s"] = time.perf_counter() - t1
        out["seconds"] = time.perf_counter() - t0
        return out

    def run(self, iterations: int, games_per_iteration: int) -> List[Dict[str, float]]:
        return [self.step(games_per_iteration) for _ in range(iterations)]


def main(argv=None) -> None:
    import argparse
    import json
    from .neural_network import ChessNet
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--iterations", type=int, default=3)
    ap.add_argument("--games", type=int, default=4096, help="games per iteration, summed over ranks")
    ap.add_argument("--sims", type=int, default=15)
    ap.add_argument("--precision", default="fp32", choices=("fp32", "tf32", "bf16"))
    ap.add_argument("--out", default=".")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-resume", action="store_true")
    a = ap.parse_args(argv)
    _lib.require_device()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(a.seed)
    net = ChessNet().cuda().eval()
    loop = TrainLoop(net, a.out, a.precision, a.sims, seed=a.seed)
    if not a.no_resume and loop.resume() and loop.rank == 0:
        print(f"resumed from {loop.latest_path}: {loop.total_games} games, {loop.training_steps} updates", flush=True)
    for _ in range(a.iterations):
        r = loop.step(a.games)
        if loop.rank == 0:
            print(json.dumps(r), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Wire / on-disk formats around the hot path, so that files produced by a
driver built on this engine stay readable by the reference's own tools:

* ``logs/training.log`` lines (trainer.py:402-406), parsed by plot_progress.py:48;
* ``models/latest.pt`` checkpoints (trainer.py:438-443 / :452-458);
* ``data/best_games.pkl`` records (trainer.py:487-495), replayed by view_best_games.py:202-213.
"""
from __future__ import annotations

from datetime import datetime
from typing import Dict, Iterable, List, Optional, Sequence, Tuple


def game_stats(results: Iterable[Tuple[Sequence, int, str]]) -> Dict[str, float]:
    """red/black wins, draws and mean sample count of a batch of (game_data, winner, reason)."""
    res = list(results)
    lengths = [len(gd) for gd, _, _ in res]
    return {"red_wins": sum(1 for _, w, _ in res if w == 1),
            "black_wins": sum(1 for _, w, _ in res if w == -1),
            "draws": sum(1 for _, w, _ in res if w == 0),
            "avg_moves": (sum(lengths) / len(lengths)) if lengths else 0.0}


def training_log_line(iteration: int, total_games: int, stats: Optional[Dict[str, float]],
                      buffer_len: int, now: Optional[datetime] = None) -> str:
    now = now or datetime.now()
    if stats:
        return (f"{now} | 轮次:{iteration} | 总局数:{total_games} | "
                f"红胜:{stats['red_wins']} 黑胜:{stats['black_wins']} 和:{stats['draws']} | "
                f"平均步数:{stats['avg_moves']:.1f} | 缓冲区:{buffer_len} | 类型:训练\n")
    return f"{now} | 轮次:{iteration} | 总局数:{total_games} | 缓冲区:{buffer_len} | 类型:训练\n"


def checkpoint_dict(network, optimizer, total_games: int, training_steps: int) -> Dict[str, object]:
    return {"model_state_dict": network.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
            "total_games": int(total_games), "training_steps": int(training_steps)}


def load_checkpoint(checkpoint: Dict[str, object], network, optimizer=None) -> Tuple[int, int]:
    network.load_state_dict(checkpoint["model_state_dict"])
    if optimizer is not None and "optimizer_state_dict" in checkpoint:
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    return int(checkpoint.get("total_games", 0)), int(checkpoint.get("training_steps", 0))


def best_game_record(game_data: List, winner: int, moves: int, game_type: str, total_games: int,
                     now: Optional[datetime] = None) -> Dict[str, object]:
    return {"timestamp": now or datetime.now(), "total_games": int(total_games),
            "game_data": game_data, "winner": int(winner), "moves": int(moves), "type": game_type}

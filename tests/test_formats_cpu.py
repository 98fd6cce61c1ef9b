"""CPU: file formats written around the engine stay readable by the reference's tools."""
import re
from datetime import datetime

import torch


def test_training_log_line_matches_plot_progress_regex():
    from chinesechessai_b200 import formats
    res = [([0] * 40, 1, "将死黑方"), ([0] * 70, 0, "超过70步判和"), ([0] * 33, -1, "困毙红方")]
    st = formats.game_stats(res)
    assert st == {"red_wins": 1, "black_wins": 1, "draws": 1, "avg_moves": (40 + 70 + 33) / 3}
    line = formats.training_log_line(7, 134, st, 9967, datetime(2025, 11, 26, 20, 31, 28, 546135))
    assert "类型:训练" in line and line.endswith("\n")
    # the exact pattern of plot_progress.py:48
    m = re.search(r'轮次:(\d+).*?总局数:(\d+).*?红胜:(\d+)\s+黑胜:(\d+)\s+和:(\d+).*?平均步数:([\d.]+)', line)
    assert m and [int(m.group(i)) for i in range(1, 6)] == [7, 134, 1, 1, 1]
    assert float(m.group(6)) == round(st["avg_moves"], 1)
    assert line.startswith("2025-11-26 20:31:28.546135 | 轮次:7 | 总局数:134 | 红胜:1 黑胜:1 和:1 | 平均步数:47.7 | ")
    old = formats.training_log_line(1, 5, None, 10)
    assert "红胜" not in old and "类型:训练" in old


def test_checkpoint_round_trip_keys():
    from chinesechessai_b200 import formats
    from chinesechessai_b200.neural_network import ChessNet
    net = ChessNet(num_channels=8)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    ck = formats.checkpoint_dict(net, opt, 1234, 56)
    assert sorted(ck) == ["model_state_dict", "optimizer_state_dict", "total_games", "training_steps"]
    net2 = ChessNet(num_channels=8)
    assert formats.load_checkpoint(ck, net2, torch.optim.Adam(net2.parameters())) == (1234, 56)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    rec = formats.best_game_record([("b", {}, 0.1)], 1, 41, "self_play", 1234)
    assert sorted(rec) == ["game_data", "moves", "timestamp", "total_games", "type", "winner"]


def test_reference_readers_accept_the_written_files(tmp_path):
    """The reference's unmodified readers (Trainer.load_model, plot_progress.parse_training_log,
    view_best_games.load_best_games) on files written with formats.py — the CPU half of
    tests/test_train_loop_gpu.py."""
    import json
    import os
    import pickle
    import subprocess
    import sys
    import numpy as np
    import pytest
    import torch
    from baseline import reference as R
    from chinesechessai_b200 import formats
    from chinesechessai_b200.neural_network import ChessNet
    if R.locate() is None:
        pytest.skip("no reference checkout")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    net = ChessNet()
    opt = torch.optim.Adam(net.parameters(), lr=0.001, weight_decay=1e-4)
    net(torch.zeros(2, 15, 10, 9))[1].sum().backward()
    opt.step()
    for d in ("models", "logs", "data"):
        os.makedirs(tmp_path / d)
    torch.save(formats.checkpoint_dict(net, opt, 96, 1), tmp_path / "models" / "latest.pt")
    st = {"red_wins": 3, "black_wins": 1, "draws": 92, "avg_moves": 69.5}
    with open(tmp_path / "logs" / "training.log", "w", encoding="utf-8") as f:
        f.write(formats.training_log_line(1, 96, st, 6648))
    board = np.zeros((10, 9), np.int8)
    board[0] = [-5, -4, -3, -2, -1, -2, -3, -4, -5]
    board[2, 1] = board[2, 7] = -6
    board[3, ::2] = -7
    board[6, ::2] = 7
    board[7, 1] = board[7, 7] = 6
    board[9] = [5, 4, 3, 2, 1, 2, 3, 4, 5]
    rec = formats.best_game_record([(board, {(6, 0, 5, 0): 0.75, (7, 1, 7, 2): 0.25}, 0.1)], 1, 1, "将死", 96)
    with open(tmp_path / "data" / "best_games.pkl", "wb") as f:
        pickle.dump([rec], f)
    p = subprocess.run([sys.executable, os.path.join(root, "tests", "drivers", "drive_consumers.py"),
                        "--mode", "readers"], env=R.env_for_reference(), cwd=tmp_path,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    got = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert got["total_games"] == 96 and got["training_steps"] == 1 and got["adam_steps"] == [1]
    assert got["log"] == {"rounds": [1], "avg_moves": [69.5], "red_wins": [3], "black_wins": [1],
                          "draws": [92], "total_games": [96]}
    assert got["best_games"] == 1 and got["replayed"][0]["winner"] == 1

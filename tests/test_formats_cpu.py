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

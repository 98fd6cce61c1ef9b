"""GPU: files written by the multi-GPU training loop (chinesechessai_b200/train_loop.py) are read
by the reference's own, unmodified tools — Trainer.load_model, plot_progress.parse_training_log,
view_best_games.load_best_games / list_best_games."""
import json
import os
import pickle
import subprocess
import sys

import pytest

from baseline import reference as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "drivers", "drive_consumers.py")


@pytest.mark.gpu
def test_train_loop_files_and_resume(tmp_path):
    import torch
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.train_loop import TrainLoop
    torch.manual_seed(0)
    games, sims = 48, 15
    loop = TrainLoop(ChessNet().cuda().eval(), str(tmp_path), precision="bf16", num_simulations=sims, seed=5)
    loop.short_draw = 1000        # random-init games are 70-ply draws: keep them so the file is not empty
    r1, r2 = loop.run(2, games)
    assert r1["games"] == games and r1["red_wins"] + r1["black_wins"] + r1["draws"] == games
    assert r1["samples"] == r1["plies"] and r1["loss"] == r1["loss"]          # a number, not NaN
    assert loop.total_games == 2 * games and loop.training_steps == 2 * min(50, r1["samples"] // 64)
    assert r1["best_games"] > 0
    lines = open(loop.log_path, encoding="utf-8").read().splitlines()
    assert len(lines) == 2 and "轮次:2" in lines[1] and f"总局数:{2 * games}" in lines[1]
    recs = pickle.load(open(loop.best_games_path, "rb"))
    assert len(recs) == r1["best_games"] + r2["best_games"]
    board0, probs0, reward0 = recs[0]["game_data"][0]
    assert type(probs0) is dict and board0.shape == (10, 9) and isinstance(reward0, float)
    assert recs[0]["moves"] == len(recs[0]["game_data"])
    # resume: a fresh loop picks up counters, weights and Adam state
    net2 = ChessNet().cuda().eval()
    loop2 = TrainLoop(net2, str(tmp_path), precision="bf16", num_simulations=sims, seed=5)
    assert loop2.resume() and loop2.total_games == 2 * games and loop2.training_steps == loop.training_steps
    for a, b in zip(loop.network.state_dict().values(), net2.state_dict().values()):
        assert torch.equal(a, b)
    steps = {int(v["step"]) for v in loop2.optimizer.state_dict()["state"].values()}
    assert steps == {loop.training_steps}

    if R.locate() is None:
        pytest.skip("no reference checkout: the reference's readers were not run")
    p = subprocess.run([sys.executable, DRIVER, "--mode", "readers"], env=R.env_for_reference(),
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-4000:]
    got = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    ref = os.path.realpath(R.locate())
    for m in ("config", "chess_env", "trainer", "plot_progress", "view_best_games"):
        assert os.path.realpath(got["modules"][m]) == ref, m            # the stock modules, not the shims
    assert got["total_games"] == 2 * games and got["training_steps"] == loop.training_steps
    assert got["adam_steps"] == [loop.training_steps]
    assert got["log"]["rounds"] == [1, 2] and got["log"]["total_games"] == [games, 2 * games]
    assert got["log"]["draws"] == [r1["draws"], r2["draws"]]
    assert got["log"]["avg_moves"] == [round(r1["avg_moves"], 1), round(r2["avg_moves"], 1)]
    assert got["best_games"] == len(recs) and len(got["replayed"]) == min(4, len(recs))

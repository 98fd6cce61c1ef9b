"""The reference's UNCHANGED consumer modules on top of the drop-in engine (north_star: "Python
keeps the ChessEnv and MCTS call surface, so trainer.py, evaluate.py and the visualizer run
unchanged"; SURVEY.md §7 test-plan item 5).

The three shim files of INTEGRATION.md Option A (integration/shims/) shadow chess_env /
self_play / neural_network; trainer.py, evaluate.py, compare_models.py and config.py are imported
from the reference checkout itself (baseline/_ref on the GPU box, see baseline/reference.py).
Skipped when no reference checkout is available."""
import json
import os
import re
import subprocess
import sys

import pytest

from baseline import reference as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "drivers", "drive_consumers.py")
needs_ref = pytest.mark.skipif(R.locate() is None, reason="no reference checkout (baseline/_ref, XQ_REFERENCE)")


def _last_json(text: str) -> dict:
    lines = [ln for ln in text.splitlines() if ln.startswith("{")]
    assert lines, text[-2000:]
    return json.loads(lines[-1])


@needs_ref
def test_shims_shadow_exactly_the_three_replaced_modules(tmp_path):
    """CPU: name resolution only — every name the consumers import from the replaced modules
    exists, and nothing else of the reference is shadowed."""
    code = (
        "import os, json, contextlib, io\n"
        "with contextlib.redirect_stdout(io.StringIO()):\n"
        "    import config, chess_env, self_play, neural_network, trainer, evaluate, compare_models\n"
        "    from self_play import MCTS, self_play_game, parallel_self_play, InterruptedWithResults, test_self_play\n"
        "    from neural_network import ChessNet, ResidualBlock, test_network\n"
        "mods = dict(config=config, chess_env=chess_env, self_play=self_play, neural_network=neural_network,\n"
        "            trainer=trainer, evaluate=evaluate, compare_models=compare_models)\n"
        "print(json.dumps({k: os.path.dirname(os.path.abspath(m.__file__)) for k, m in mods.items()}\n"
        "      | {'trainer_net': trainer.ChessNet.__module__, 'match_env': compare_models.ChineseChess.__module__,\n"
        "         'eval_game': evaluate.self_play_game.__module__}))\n")
    p = subprocess.run([sys.executable, "-c", code], env=R.env_for_shims(), cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    got = _last_json(p.stdout)
    ref = os.path.realpath(R.locate())
    for m in ("chess_env", "self_play", "neural_network"):
        assert os.path.realpath(got[m]) == os.path.realpath(R.SHIMS)
    for m in ("config", "trainer", "evaluate", "compare_models"):
        assert os.path.realpath(got[m]) == ref
    assert got["trainer_net"] == "chinesechessai_b200.neural_network"
    assert got["match_env"] == "chinesechessai_b200.chess_env"
    assert got["eval_game"] == "chinesechessai_b200.self_play"


@needs_ref
@pytest.mark.gpu
def test_unchanged_trainer_evaluate_compare_models_run_on_the_engine(tmp_path):
    """GPU: Trainer.collect_self_play_data -> parallel_self_play -> train_network, _log_progress,
    save_model, evaluate.evaluate_model and compare_models.play_match (trainer.py:147-362,
    :395-449; evaluate.py:13-132; compare_models.py:13-92), none of them modified."""
    games = 12
    p = subprocess.run([sys.executable, DRIVER, "--mode", "consumers", "--games", str(games),
                        "--eval-games", "2", "--match-games", "2"],
                       env=R.env_for_shims(), cwd=tmp_path, capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stderr[-4000:]
    got = _last_json(p.stdout)
    ref = os.path.realpath(R.locate())
    assert os.path.realpath(got["modules"]["trainer"]) == ref
    assert os.path.realpath(got["modules"]["config"]) == ref
    assert os.path.realpath(got["modules"]["self_play"]) == os.path.realpath(R.SHIMS)
    assert got["engine"] == "chinesechessai_b200.chess_env" and got["device"] == "cuda"
    c = got["collect"]
    assert c["red_wins"] + c["black_wins"] + c["draws"] == games and got["total_games"] == games
    assert got["buffer"] == round(c["avg_moves"] * games) and got["buffer"] >= games
    s = got["sample"]
    assert s["board_shape"] == [10, 9] and s["board_dtype"] == "int8" and s["reward_type"] == "float"
    assert s["n_probs"] == 44 and abs(s["probs_sum"] - 1.0) < 1e-9      # first sample = initial position
    assert got["train_loss"] == got["train_loss"] and got["train_loss"] >= 0.0
    assert got["training_steps"] == min(50, got["buffer"] // 64)
    # the line Trainer._log_progress wrote parses with the reference's own plot regex
    # (plot_progress.py:48)
    assert re.search(r"轮次:(\d+).*总局数:(\d+).*缓冲区:(\d+)", got["log_line"])
    e = got["evaluate"]
    assert e["red_wins"] + e["black_wins"] + e["draws"] == 2 and 1 <= e["min_moves"] <= e["max_moves"] <= 70
    m = got["play_match"]
    assert m["model1_wins"] + m["model2_wins"] + m["draws"] == 2 and 0 < m["avg_moves"] <= 100

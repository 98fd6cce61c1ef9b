# INTEGRATION.md Option A: drop-in for the reference's self_play.py.
from chinesechessai_b200.self_play import (InterruptedWithResults, MCTS,  # noqa: F401
                                           parallel_self_play, self_play_game, test_self_play)

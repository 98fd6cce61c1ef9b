# INTEGRATION.md Option A: drop-in for the reference's chess_env.py (768 lines of Python rules).
from chinesechessai_b200.chess_env import ChineseChess  # noqa: F401

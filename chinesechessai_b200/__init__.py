"""chinesechessai_b200 — B200-native (sm_100a) Xiangqi self-play hot path.

Drop-in for the rules engine (chess_env.py) and MCTS loop (self_play.py) of
hpy666666/ChineseChessAI.  All compute runs in hand-written CUDA kernels behind
the C ABI of include/xq_b200.h (libxq_b200.so); there is no CPU fallback.
"""
__version__ = "0.1.0"

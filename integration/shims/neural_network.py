# INTEGRATION.md Option A: drop-in for the reference's neural_network.py (same state_dict keys).
from chinesechessai_b200.neural_network import ChessNet, ResidualBlock, test_network  # noqa: F401

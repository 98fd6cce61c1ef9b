"""Constants of the reference's config.py that are part of the hot path's contract
(config.py:9-10,62-74): the 70-ply cap, the default simulation count, board size, piece codes."""
MAX_MOVES = 70
MCTS_SIMULATIONS = 50
NUM_WORKERS = 4
BOARD_SIZE = 10
BOARD_WIDTH = 9
PIECES = {
    "EMPTY": 0,
    "R_KING": 1, "R_ADVISOR": 2, "R_BISHOP": 3, "R_KNIGHT": 4, "R_ROOK": 5, "R_CANNON": 6, "R_PAWN": 7,
    "B_KING": -1, "B_ADVISOR": -2, "B_BISHOP": -3, "B_KNIGHT": -4, "B_ROOK": -5, "B_CANNON": -6,
    "B_PAWN": -7,
}

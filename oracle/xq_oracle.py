"""ctypes binding of the CPU oracle (oracle/xq_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``chinesechessai_b200/`` may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libxq_oracle.so")

NSQ = 90
MAX_MOVES = 128
HIST_CAP = 1024
WINNER_NONE = 2

REASON_NONE, REASON_KING_CAPTURE, REASON_CHECKMATE, REASON_REPETITION, REASON_FIFTY, \
    REASON_STALEMATE, REASON_PERPETUAL_CHECK, REASON_PERPETUAL_CHASE, REASON_MOVE_CAP = range(9)


class State(C.Structure):
    _fields_ = [
        ("board", C.c_int8 * NSQ),
        ("pad_", C.c_int8 * 6),
        ("player", C.c_int32),
        ("move_count", C.c_int32),
        ("winner", C.c_int32),
        ("reason", C.c_int32),
        ("red_king", C.c_int32),
        ("black_king", C.c_int32),
        ("no_capture", C.c_int32),
        ("consecutive_checks", C.c_int32),
        ("pos_len", C.c_int32),
        ("check_len", C.c_int32),
        ("overflow", C.c_int32),
        ("pad2_", C.c_int32),
        ("pos_hist", C.c_uint64 * HIST_CAP),
        ("check_hist", C.c_uint8 * HIST_CAP),
    ]


class StepResult(C.Structure):
    _fields_ = [("reward", C.c_double), ("reward_is_int", C.c_int32), ("done", C.c_int32)]


class PlayoutResult(C.Structure):
    _fields_ = [
        ("plies", C.c_int32),
        ("winner", C.c_int32),
        ("reason", C.c_int32),
        ("max_legal", C.c_int32),
        ("reward_sum", C.c_double),
        ("digest", C.c_uint64),
        ("final_hash", C.c_uint64),
    ]


PLAYOUT_RESULT_DTYPE = np.dtype(
    [("plies", "<i4"), ("winner", "<i4"), ("reason", "<i4"), ("max_legal", "<i4"),
     ("reward_sum", "<f8"), ("digest", "<u8"), ("final_hash", "<u8")])

EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_int8), C.POINTER(C.c_int32),
                      C.POINTER(C.c_int16), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                      C.POINTER(C.c_double))

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    src = [os.path.join(_HERE, "xq_oracle.c"), os.path.join(_HERE, "xq_oracle.h")]
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        P = C.POINTER
        L.xqo_reset.argtypes = [P(State)]
        L.xqo_position_hash.argtypes = [C.c_void_p, C.c_int]
        L.xqo_position_hash.restype = C.c_uint64
        L.xqo_legal_moves.argtypes = [P(State), C.c_void_p]
        L.xqo_legal_moves.restype = C.c_int
        L.xqo_pseudo_count.argtypes = [P(State)]
        L.xqo_pseudo_count.restype = C.c_int
        L.xqo_is_in_check.argtypes = [P(State), C.c_int]
        L.xqo_is_in_check.restype = C.c_int
        L.xqo_kings_facing.argtypes = [P(State)]
        L.xqo_kings_facing.restype = C.c_int
        L.xqo_make_move.argtypes = [P(State), C.c_int, P(StepResult)]
        L.xqo_philox4x32.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        L.xqo_pick_move.argtypes = [P(State), C.c_void_p, C.c_int, C.c_uint64, C.c_uint32,
                                    C.c_uint32, C.c_int]
        L.xqo_pick_move.restype = C.c_int
        L.xqo_playout.argtypes = [P(State), C.c_uint64, C.c_uint32, C.c_int, C.c_int,
                                  P(PlayoutResult)] + [C.c_void_p] * 6
        L.xqo_playout_many.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p]
        L.xqo_playout_many.restype = C.c_int64
        L.xqo_sample_move.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint32, C.c_uint32]
        L.xqo_sample_move.restype = C.c_int
        L.xqo_encode_board.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.xqo_logits_to_priors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.xqo_mcts_search.argtypes = [P(State), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
        L.xqo_mcts_search.restype = C.c_int
        L.xqo_hash_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def pack(move: Sequence[int]) -> int:
    fr, fc, tr, tc = move
    return (fr * 9 + fc) * 90 + tr * 9 + tc


def unpack(m: int) -> Tuple[int, int, int, int]:
    f, t = divmod(int(m), 90)
    return (f // 9, f % 9, t // 9, t % 9)


class Env:
    """One oracle board with the field names of the reference's ChineseChess."""

    def __init__(self):
        self.s = State()
        lib().xqo_reset(C.byref(self.s))

    # -- state access -------------------------------------------------------
    @property
    def board(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.s.board).reshape(10, 9)

    def set_board(self, b) -> None:
        np.ctypeslib.as_array(self.s.board)[:] = np.asarray(b, dtype=np.int8).reshape(90)

    @staticmethod
    def _sq(pos) -> int:
        return -1 if pos is None else int(pos[0]) * 9 + int(pos[1])

    @staticmethod
    def _pos(sq: int):
        return None if sq < 0 else (sq // 9, sq % 9)

    def load(self, board, player, move_count=0, winner=None, red_king=None, black_king=None,
             no_capture=0, consecutive_checks=0, check_history=(), position_history=()):
        self.set_board(board)
        s = self.s
        s.player, s.move_count = int(player), int(move_count)
        s.winner = WINNER_NONE if winner is None else int(winner)
        s.reason = REASON_NONE
        s.red_king, s.black_king = self._sq(red_king), self._sq(black_king)
        s.no_capture, s.consecutive_checks = int(no_capture), int(consecutive_checks)
        s.check_len = len(check_history)
        for i, v in enumerate(check_history):
            s.check_hist[i] = 1 if v else 0
        s.pos_len = len(position_history)
        for i, v in enumerate(position_history):
            s.pos_hist[i] = int(v)
        return self

    def clone(self) -> "Env":
        e = Env.__new__(Env)
        e.s = State()
        C.memmove(C.byref(e.s), C.byref(self.s), C.sizeof(State))
        return e

    @property
    def current_player(self) -> int:
        return self.s.player

    @property
    def winner(self):
        return None if self.s.winner == WINNER_NONE else self.s.winner

    @property
    def red_king_pos(self):
        return self._pos(self.s.red_king)

    @property
    def black_king_pos(self):
        return self._pos(self.s.black_king)

    @property
    def check_history(self) -> List[bool]:
        return [bool(self.s.check_hist[i]) for i in range(self.s.check_len)]

    @property
    def position_history(self) -> List[int]:
        return [int(self.s.pos_hist[i]) for i in range(self.s.pos_len)]

    # -- rules --------------------------------------------------------------
    def legal_moves_packed(self) -> np.ndarray:
        buf = np.zeros(MAX_MOVES, dtype=np.int16)
        n = lib().xqo_legal_moves(C.byref(self.s), buf.ctypes.data)
        return buf[:n].copy()

    def get_legal_moves(self) -> List[Tuple[int, int, int, int]]:
        return [unpack(m) for m in self.legal_moves_packed()]

    def pseudo_count(self) -> int:
        return lib().xqo_pseudo_count(C.byref(self.s))

    def is_in_check(self, player: int) -> bool:
        return bool(lib().xqo_is_in_check(C.byref(self.s), int(player)))

    def kings_facing(self) -> bool:
        return bool(lib().xqo_kings_facing(C.byref(self.s)))

    def make_move(self, move) -> Tuple[float, bool, bool]:
        """Returns (reward, reward_is_int, done)."""
        m = move if isinstance(move, (int, np.integer)) else pack(move)
        r = StepResult()
        lib().xqo_make_move(C.byref(self.s), int(m), C.byref(r))
        return r.reward, bool(r.reward_is_int), bool(r.done)

    def position_hash(self, player: Optional[int] = None) -> int:
        p = self.s.player if player is None else player
        return int(lib().xqo_position_hash(C.addressof(self.s.board), int(p)))

    def playout(self, seed: int, game_id: int, max_plies: int = 70, capture_bias: int = 0,
                trace: bool = False):
        res = PlayoutResult()
        if trace:
            tm = np.zeros((max_plies, MAX_MOVES), np.int16)
            tn = np.zeros(max_plies, np.int16)
            tp = np.zeros(max_plies, np.int16)
            tr = np.zeros(max_plies, np.float64)
            tf = np.zeros(max_plies, np.uint8)
            tb = np.zeros((max_plies, NSQ), np.int8)
            ptrs = [a.ctypes.data for a in (tm, tn, tp, tr, tf, tb)]
        else:
            ptrs = [None] * 6
        lib().xqo_playout(C.byref(self.s), seed, game_id, max_plies, capture_bias, C.byref(res),
                          *ptrs)
        if trace:
            return res, dict(moves=tm, n=tn, pick=tp, reward=tr, flags=tf, boards=tb)
        return res


def playout_many(n_games: int, seed: int, first_game_id: int = 0, max_plies: int = 70,
                 capture_bias: int = 0, n_threads: int = 1):
    res = np.zeros(n_games, dtype=PLAYOUT_RESULT_DTYPE)
    total = lib().xqo_playout_many(n_games, first_game_id, seed, max_plies, capture_bias,
                                   n_threads, res.ctypes.data)
    return int(total), res


def philox(c0: int, c1: int, k0: int, k1: int) -> np.ndarray:
    out = np.zeros(4, np.uint32)
    lib().xqo_philox4x32(c0, c1, 0, 0, k0, k1, out.ctypes.data)
    return out


def sample_move(visits, temperature: float, seed: int, game_id: int, ply: int) -> int:
    v = np.ascontiguousarray(visits, dtype=np.int32)
    return int(lib().xqo_sample_move(v.ctypes.data, len(v), float(temperature), seed, game_id, ply))


def encode_board(board, player: int) -> np.ndarray:
    b = np.ascontiguousarray(np.asarray(board, dtype=np.int8).reshape(90))
    out = np.zeros((15, 10, 9), np.float32)
    lib().xqo_encode_board(b.ctypes.data, int(player), out.ctypes.data)
    return out


def logits_to_priors(logits: np.ndarray, moves_packed: np.ndarray) -> np.ndarray:
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    mv = np.ascontiguousarray(moves_packed, dtype=np.int16)
    out = np.zeros(len(mv), np.float32)
    lib().xqo_logits_to_priors(lg.ctypes.data, mv.ctypes.data, len(mv), out.ctypes.data)
    return out


def hash_eval(boards: np.ndarray, players: np.ndarray, moves: np.ndarray, n_moves: np.ndarray,
              flat: bool = False):
    """Built-in deterministic evaluator on numpy batches -> (priors f32[n,128], values f64[n])."""
    n = len(players)
    boards = np.ascontiguousarray(boards, np.int8).reshape(n, 90)
    players = np.ascontiguousarray(players, np.int32)
    moves = np.ascontiguousarray(moves, np.int16).reshape(n, MAX_MOVES)
    n_moves = np.ascontiguousarray(n_moves, np.int32)
    pri = np.zeros((n, MAX_MOVES), np.float32)
    val = np.zeros(n, np.float64)
    mode = C.c_int(1 if flat else 0)
    lib().xqo_hash_eval(C.addressof(mode), n, boards.ctypes.data, players.ctypes.data,
                        moves.ctypes.data, n_moves.ctypes.data, pri.ctypes.data, val.ctypes.data)
    return pri, val


def mcts_search(env: Env, num_simulations: int,
                evaluator: Optional[Callable] = None, flat: bool = False):
    """self_play.py:89-154 on the oracle.  ``evaluator(boards i8[n,90], players i32[n],
    moves i16[n,128], n_moves i32[n]) -> (priors f32[n,128], values f64[n])``; None = the
    built-in hash evaluator.  Returns (moves_packed, visits, stats)."""
    rm = np.zeros(MAX_MOVES, np.int16)
    rv = np.zeros(MAX_MOVES, np.int32)
    st = np.zeros(4, np.int64)
    L = lib()
    if evaluator is None:
        mode = C.c_int(1 if flat else 0)
        fn = C.cast(L.xqo_hash_eval, C.c_void_p)
        n = L.xqo_mcts_search(C.byref(env.s), num_simulations, fn, C.addressof(mode),
                              rm.ctypes.data, rv.ctypes.data, st.ctypes.data)
    else:
        def _cb(ctx, nl, boards, players, moves, n_moves, priors, values):
            b = np.ctypeslib.as_array(boards, (nl, 90))
            p = np.ctypeslib.as_array(players, (nl,))
            m = np.ctypeslib.as_array(moves, (nl, MAX_MOVES))
            k = np.ctypeslib.as_array(n_moves, (nl,))
            pr, va = evaluator(b, p, m, k)
            np.ctypeslib.as_array(priors, (nl, MAX_MOVES))[:] = pr
            np.ctypeslib.as_array(values, (nl,))[:] = va
        cb = EVAL_FN(_cb)
        n = L.xqo_mcts_search(C.byref(env.s), num_simulations, C.cast(cb, C.c_void_p), None,
                              rm.ctypes.data, rv.ctypes.data, st.ctypes.data)
    return rm[:n].copy(), rv[:n].copy(), st

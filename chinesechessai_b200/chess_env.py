"""Drop-in ``ChineseChess`` (reference: chess_env.py) over the CUDA rules engine.

Same class surface as the reference: public mutable attributes (``board``, ``current_player``,
``move_count``, ``winner``, ``end_reason``, king caches, counters, history lists) live on the
host exactly as callers and the reference's scripts poke them; every rules computation
(get_legal_moves, make_move, the check queries) marshals that state to the GPU and runs the same
kernels as the batched engine through the C ABI.  There is no CPU rules fallback.

This one-board view costs a launch + small copies per call; throughput work uses
``engine.BoardBatch`` / ``mcts.BatchedMCTS`` with state resident in HBM.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BOARD_STRIDE, MAX_MOVES, META_DTYPE, check
from .config import BOARD_SIZE, BOARD_WIDTH
from .engine import _ptr, _stream, pack_move, unpack_move

Move = Tuple[int, int, int, int]

_RED, _BLACK = "红方", "黑方"
_MOVE_TUPLES: List[Move] = []


def _move_tuples() -> List[Move]:
    """(from_r, from_c, to_r, to_c) of every packed move from*90+to, built once."""
    if not _MOVE_TUPLES:
        _MOVE_TUPLES.extend(unpack_move(m) for m in range(_lib.POLICY))
    return _MOVE_TUPLES


def format_end_reason(reason: int, side_to_move: int, move_count: int) -> Optional[str]:
    """The reference's end_reason strings (chess_env.py:297,359,366,373,381,389,404).
    ``side_to_move`` is current_player AFTER the move (the loser for mate/stalemate/perpetual
    check; the mover is its opposite)."""
    stm = _RED if side_to_move == 1 else _BLACK
    mover = _BLACK if side_to_move == 1 else _RED
    return {
        _lib.REASON_NONE: None,
        _lib.REASON_KING_CAPTURE: f"{mover}吃掉对方将帅",
        _lib.REASON_CHECKMATE: f"将死{stm}",
        _lib.REASON_REPETITION: "三次重复局面判和",
        _lib.REASON_FIFTY: "50回合无吃子判和",
        _lib.REASON_STALEMATE: f"困毙{stm}",
        _lib.REASON_PERPETUAL_CHECK: f"长将判负({stm})",
        _lib.REASON_PERPETUAL_CHASE: f"长捉判负({stm})",
        _lib.REASON_MOVE_CAP: f"超过{move_count}步判和",
    }[int(reason)]


class _Device1:
    """Device + pinned staging for ONE board (shared by all envs of a process).

    Everything a call exchanges with the GPU lives in one byte arena, mirrored in pinned host
    memory: a rules call is ONE host->device copy, ONE kernel and ONE device->host copy (round 1
    paid a separate synchronising copy per field — five per make_move — which is what the
    reference's callers feel: evaluate.py / compare_models.py make ~1,400 such calls per match).

    Arena layout (bytes): board 0:96 | meta 96:128 | move 128:130 | reward 136:144 | flags 144 |
    query 152:156 | key 160:168 | n_moves 168:170 | moves 176:432 | position history 432:..."""
    _inst = None
    O_BOARD, O_META, O_MOVE, O_REWARD, O_FLAGS, O_Q, O_KEY, O_NMOVES, O_MOVES, O_HIST = \
        0, 96, 128, 136, 144, 152, 160, 168, 176, 432

    def __init__(self):
        self.lib = _lib.load()
        _lib.require_device()
        if not torch.cuda.is_available():
            raise _lib.XqError("torch sees no CUDA device; ChineseChess has no CPU fallback")
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.hist_cap = 0
        self._alloc(256)
        init_b = torch.zeros((1, BOARD_STRIDE), dtype=torch.int8, device=self.dev)
        init_m = torch.zeros((1, 32), dtype=torch.uint8, device=self.dev)
        check(self.lib.xq_reset(_ptr(init_b), _ptr(init_m), 1, _stream()))
        self.init_board = init_b[0, :90].cpu().numpy().reshape(BOARD_SIZE, BOARD_WIDTH).copy()

    def _alloc(self, hist_cap: int) -> None:
        self.hist_cap = hist_cap
        size = self.O_HIST + 8 * hist_cap
        self.d = torch.zeros(size, dtype=torch.uint8, device=self.dev)
        self.h = torch.zeros(size, dtype=torch.uint8).pin_memory()
        self.hn = self.h.numpy()
        base = self.d.data_ptr()
        self.p = {k: base + getattr(self, "O_" + k) for k in
                  ("BOARD", "META", "MOVE", "REWARD", "FLAGS", "Q", "KEY", "NMOVES", "MOVES", "HIST")}
        # typed host views of the pinned mirror
        self.h_board = self.hn[self.O_BOARD:self.O_BOARD + BOARD_STRIDE].view(np.int8)
        self.h_meta = self.hn[self.O_META:self.O_META + 32].view(META_DTYPE)
        self.h_move = self.hn[self.O_MOVE:self.O_MOVE + 2].view(np.int16)
        self.h_reward = self.hn[self.O_REWARD:self.O_REWARD + 8].view(np.float64)
        self.h_nmoves = self.hn[self.O_NMOVES:self.O_NMOVES + 2].view(np.int16)
        self.h_moves = self.hn[self.O_MOVES:self.O_MOVES + 2 * MAX_MOVES].view(np.int16)
        self.h_key = self.hn[self.O_KEY:self.O_KEY + 8].view(np.uint64)
        self.h_hist = self.hn[self.O_HIST:].view(np.uint64)

    @classmethod
    def get(cls) -> "_Device1":
        if cls._inst is None or cls._inst.dev.index != torch.cuda.current_device():
            cls._inst = cls()
        return cls._inst

    def need_hist(self, n: int) -> None:
        if n + 1 > self.hist_cap:
            self._alloc(max(2 * self.hist_cap, n + 64))

    def up(self, lo: int, hi: int) -> None:
        self.d[lo:hi].copy_(self.h[lo:hi], non_blocking=True)

    def down(self, lo: int, hi: int) -> None:
        self.h[lo:hi].copy_(self.d[lo:hi], non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()


class _LazyThreats(list):
    """One entry of ``chase_history`` (chess_env.py:344-345: ``_get_threatened_pieces`` of the side
    that has just moved), computed on first use.

    The reference fills it on every ply and never reads it (its only consumer returns False first,
    :674), which is ~70 % of its ``make_move`` time.  Here the entry keeps a snapshot of the
    position after the move and asks the engine only if somebody looks: the mover's legal moves
    (same generators, same suicide filter, geometry of the mover) that capture a non-king piece,
    in scan order, as ``((r, c), (to_r, to_c))`` pairs.  ``_is_protected`` can never be true in
    the reference — a piece cannot "move" onto a piece of its own side, :116 — so every such
    capture is listed, exactly as the reference lists it."""
    __slots__ = ("_snap",)

    def _fill(self) -> None:
        snap = getattr(self, "_snap", None)
        if snap is None:
            return
        self._snap = None
        board, mover, red, black = snap
        env = ChineseChess.__new__(ChineseChess)
        env.board, env.current_player = board, mover
        env.red_king_pos, env.black_king_pos = red, black
        env.winner, env.move_count, env.no_capture_count, env.consecutive_checks = None, 0, 0, 0
        env.check_history, env.position_history = [], []
        for fr, fc, tr, tc in env.get_legal_moves():
            target = int(board[tr, tc])
            if target * mover < 0 and abs(target) != 1:
                list.append(self, ((fr, fc), (tr, tc)))

    def _filled(name):  # noqa: N805
        base = getattr(list, name)

        def method(self, *a, **k):
            self._fill()
            return base(self, *a, **k)
        method.__name__ = name
        return method

    for _n in ("__getitem__", "__iter__", "__len__", "__contains__", "__eq__", "__ne__", "__repr__",
               "__reversed__", "__add__", "__setitem__", "__delitem__", "append", "extend", "index", "count",
               "copy", "pop", "insert", "remove", "sort", "reverse"):
        locals()[_n] = _filled(_n)
    del _n, _filled

    def __bool__(self) -> bool:
        return self.__len__() > 0

    def __reduce__(self):
        self._fill()
        return (list, (list(self),))


class ChineseChess:
    def __init__(self) -> None:
        self.reset()

    # -- chess_env.py:14-67 ----------------------------------------------------------------
    def reset(self):
        dev = _Device1.get()
        self.board = dev.init_board.copy()
        self.position_history: List[int] = []
        self.no_capture_count = 0
        self.check_history: List[bool] = []
        self.chase_history: List[list] = []  # entries are computed on first use (_LazyThreats)
        self.consecutive_checks = 0
        self.red_king_pos: Optional[Tuple[int, int]] = (9, 4)
        self.black_king_pos: Optional[Tuple[int, int]] = (0, 4)
        self.current_player = 1
        self.move_count = 0
        self.winner = None
        self.end_reason = None
        return self.get_state()

    def get_state(self):
        return self.board.copy(), self.current_player

    # -- host <-> device marshalling ---------------------------------------------------------
    @staticmethod
    def _sq(pos) -> int:
        return -1 if pos is None else int(pos[0]) * 9 + int(pos[1])

    def _stage(self, dev: _Device1, with_hist: bool = False) -> int:
        """Write this env into the pinned arena; returns the end offset of what has to go up."""
        if with_hist:
            dev.need_hist(len(self.position_history))
        dev.h_board[:90] = np.asarray(self.board, dtype=np.int8).reshape(90)
        dev.h_board[90:] = 0
        m = dev.h_meta
        m[:] = 0
        m["player"] = 1 if self.current_player == 1 else -1
        m["winner"] = _lib.WINNER_NONE if self.winner is None else int(self.winner)
        m["red_king"], m["black_king"] = self._sq(self.red_king_pos), self._sq(self.black_king_pos)
        m["move_count"], m["no_capture"] = int(self.move_count), int(self.no_capture_count)
        m["consecutive_checks"] = int(self.consecutive_checks)
        ck = self.check_history
        m["check_len"] = len(ck)
        m["check_bits"] = sum((1 << i) for i, v in enumerate(ck[-32:][::-1]) if v)
        n_hist = len(self.position_history) if with_hist else 0
        m["hist_len"] = n_hist
        if n_hist:
            dev.h_hist[:n_hist] = np.array(self.position_history, dtype=np.uint64)
        return dev.O_HIST + 8 * n_hist if with_hist else dev.O_META + 32

    # -- chess_env.py:76-121 -----------------------------------------------------------------
    def get_legal_moves(self) -> List[Move]:
        dev = _Device1.get()
        self._stage(dev)
        dev.up(0, dev.O_META + 32)
        p = dev.p
        check(dev.lib.xq_legal_moves(p["BOARD"], p["META"], p["MOVES"], p["NMOVES"], None, 1, _stream()))
        dev.down(dev.O_META, dev.O_MOVES + 2 * MAX_MOVES)
        if int(dev.h_meta["flags"][0]) & _lib.F_OVERFLOW:
            raise _lib.XqError("position exceeds the engine's capacity (>128 legal moves or >256 "
                               "pseudo-legal candidates)")
        n = int(dev.h_nmoves[0])
        tup = _move_tuples()
        return [tup[m] for m in dev.h_moves[:n].tolist()]

    # -- chess_env.py:253-406 ----------------------------------------------------------------
    def make_move(self, move: Move):
        fr, fc, tr, tc = (int(x) for x in move)
        if not (0 <= fr < BOARD_SIZE and 0 <= tr < BOARD_SIZE and 0 <= fc < BOARD_WIDTH and 0 <= tc < BOARD_WIDTH):
            raise IndexError(f"move {move} is off the 10x9 board")  # numpy raises here too (>= size)
        dev = _Device1.get()
        end = self._stage(dev, with_hist=True)
        dev.h_move[0] = pack_move((fr, fc, tr, tc))
        dev.up(0, end)
        p = dev.p
        check(dev.lib.xq_step(p["BOARD"], p["META"], p["HIST"], dev.hist_cap, p["MOVE"], p["REWARD"],
                              p["FLAGS"], None, None, 1, _stream()))
        n_before = len(self.position_history)
        dev.down(0, dev.O_HIST + 8 * (n_before + 1))
        m = dev.h_meta[0]
        if int(m["flags"]) & _lib.F_OVERFLOW:
            raise _lib.XqError("position exceeds the engine's capacity (>128 legal moves, >256 "
                               "pseudo-legal candidates or a full position history)")
        flags = int(dev.hn[dev.O_FLAGS])
        reward = float(dev.h_reward[0])
        self.board = dev.h_board[:90].reshape(BOARD_SIZE, BOARD_WIDTH).copy()
        pos = lambda s: None if s < 0 else (int(s) // 9, int(s) % 9)
        self.red_king_pos, self.black_king_pos = pos(m["red_king"]), pos(m["black_king"])
        self.no_capture_count = int(m["no_capture"])
        self.consecutive_checks = int(m["consecutive_checks"])
        n_hist = int(m["hist_len"])
        if n_hist > n_before:
            self.position_history.append(int(dev.h_hist[n_hist - 1]))
        self.check_history.append(bool(int(m["check_bits"]) & 1))
        threats = _LazyThreats()   # :344-345, evaluated for the mover before the side switch
        threats._snap = (self.board.copy(), -int(m["player"]), self.red_king_pos, self.black_king_pos)
        self.chase_history.append(threats)
        self.current_player = int(m["player"])
        self.move_count = int(m["move_count"])
        done = bool(flags & 1)
        if done:
            self.winner = None if m["winner"] == _lib.WINNER_NONE else int(m["winner"])
            self.end_reason = format_end_reason(int(m["reason"]), self.current_player, self.move_count)
        if flags & 2:
            reward = int(reward)  # the reference returns a Python int on these paths
        return self.get_state(), reward, done

    # -- private helpers the reference's own scripts call -------------------------------------
    def _query(self):
        dev = _Device1.get()
        self._stage(dev)
        dev.up(0, dev.O_META + 32)
        check(dev.lib.xq_query_checks(dev.p["BOARD"], dev.p["META"], dev.p["Q"], 1, _stream()))
        dev.down(dev.O_Q, dev.O_Q + 4)
        return dev.hn[dev.O_Q:dev.O_Q + 4].tolist()

    def _is_in_check(self, player) -> bool:
        q = self._query()
        return bool(q[0] if player == 1 else q[1])

    def _are_kings_facing(self) -> bool:
        return bool(self._query()[2])

    def _get_position_hash(self) -> int:
        dev = _Device1.get()
        self._stage(dev)
        dev.up(0, dev.O_META + 32)
        check(dev.lib.xq_position_hash(dev.p["BOARD"], dev.p["META"], dev.p["KEY"], 1, _stream()))
        dev.down(dev.O_KEY, dev.O_KEY + 8)
        return int(dev.h_key[0])

    def _check_draw_by_repetition(self) -> bool:  # :598-605
        return self.position_history.count(self._get_position_hash()) >= 3

    def _check_draw_by_fifty_moves(self) -> bool:  # :607-612
        return self.no_capture_count >= 100

    def _check_checkmate(self) -> bool:  # :614-628
        return len(self.get_legal_moves()) == 0 and self._is_in_check(self.current_player)

    def _check_stalemate(self) -> bool:  # :630-644
        return len(self.get_legal_moves()) == 0 and not self._is_in_check(self.current_player)

    def _check_perpetual_check(self) -> bool:  # :646-662
        return len(self.check_history) >= 12 and sum(1 for c in self.check_history[-12:] if c) >= 10

    def _check_perpetual_chase(self) -> bool:  # :664-674 — disabled in the reference
        return False

    # -- chess_env.py:408-429 ------------------------------------------------------------------
    def render(self) -> None:
        names = {0: "·", 1: "帅", 2: "士", 3: "相", 4: "马", 5: "车", 6: "炮", 7: "兵",
                 -1: "将", -2: "士", -3: "象", -4: "马", -5: "车", -6: "炮", -7: "卒"}
        print("\n  " + "".join(f"{i} " for i in range(BOARD_WIDTH)))
        for r in range(BOARD_SIZE):
            print(f"{r} " + "".join(names[int(self.board[r, c])] + " " for c in range(BOARD_WIDTH)))
        print(f"\n当前: {_RED if self.current_player == 1 else _BLACK}")
        print(f"步数: {self.move_count}")

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
    """Device + pinned staging buffers for a single board (shared by all envs of a process)."""
    _inst = None

    def __init__(self):
        self.lib = _lib.load()
        _lib.require_device()
        if not torch.cuda.is_available():
            raise _lib.XqError("torch sees no CUDA device; ChineseChess has no CPU fallback")
        d = self.dev = torch.device("cuda", torch.cuda.current_device())
        self.board = torch.zeros((1, BOARD_STRIDE), dtype=torch.int8, device=d)
        self.meta = torch.zeros((1, 32), dtype=torch.uint8, device=d)
        self.moves = torch.zeros((1, MAX_MOVES), dtype=torch.int16, device=d)
        self.n_moves = torch.zeros((1,), dtype=torch.int16, device=d)
        self.move = torch.zeros((1,), dtype=torch.int16, device=d)
        self.reward = torch.zeros((1,), dtype=torch.float64, device=d)
        self.flags = torch.zeros((1,), dtype=torch.uint8, device=d)
        self.q = torch.zeros((1, 4), dtype=torch.uint8, device=d)
        self.key = torch.zeros((1,), dtype=torch.int64, device=d)
        self.hist_cap = 256
        self.hist = torch.zeros((1, self.hist_cap), dtype=torch.int64, device=d)
        init_b = torch.zeros((1, BOARD_STRIDE), dtype=torch.int8, device=d)
        init_m = torch.zeros((1, 32), dtype=torch.uint8, device=d)
        check(self.lib.xq_reset(_ptr(init_b), _ptr(init_m), 1, _stream()))
        self.init_board = init_b[0, :90].cpu().numpy().reshape(BOARD_SIZE, BOARD_WIDTH).copy()

    @classmethod
    def get(cls) -> "_Device1":
        if cls._inst is None or cls._inst.dev.index != torch.cuda.current_device():
            cls._inst = cls()
        return cls._inst

    def need_hist(self, n: int) -> None:
        if n + 1 > self.hist_cap:
            self.hist_cap = max(2 * self.hist_cap, n + 64)
            self.hist = torch.zeros((1, self.hist_cap), dtype=torch.int64, device=self.dev)


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
        self.chase_history: List[list] = []  # the reference's chase scan is dead work (:345,:674)
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

    def _upload(self, dev: _Device1, with_hist: bool = False) -> None:
        b = np.zeros((1, BOARD_STRIDE), np.int8)
        b[0, :90] = np.asarray(self.board, dtype=np.int8).reshape(90)
        m = np.zeros(1, META_DTYPE)
        m["player"] = 1 if self.current_player == 1 else -1
        m["winner"] = _lib.WINNER_NONE if self.winner is None else int(self.winner)
        m["red_king"], m["black_king"] = self._sq(self.red_king_pos), self._sq(self.black_king_pos)
        m["move_count"], m["no_capture"] = int(self.move_count), int(self.no_capture_count)
        m["consecutive_checks"] = int(self.consecutive_checks)
        ck = self.check_history
        m["check_len"] = len(ck)
        m["check_bits"] = sum((1 << i) for i, v in enumerate(ck[-32:][::-1]) if v)
        m["hist_len"] = len(self.position_history) if with_hist else 0
        dev.board.copy_(torch.from_numpy(b))
        dev.meta.copy_(torch.from_numpy(m.view(np.uint8).reshape(1, 32)))
        if with_hist and self.position_history:
            dev.need_hist(len(self.position_history))
            h = np.array(self.position_history, dtype=np.uint64).view(np.int64)
            dev.hist[0, :len(h)].copy_(torch.from_numpy(h))
        elif with_hist:
            dev.need_hist(0)

    # -- chess_env.py:76-121 -----------------------------------------------------------------
    def get_legal_moves(self) -> List[Move]:
        dev = _Device1.get()
        self._upload(dev)
        check(dev.lib.xq_legal_moves(_ptr(dev.board), _ptr(dev.meta), _ptr(dev.moves),
                                     _ptr(dev.n_moves), None, 1, _stream()))
        n = int(dev.n_moves[0])
        if int(dev.meta[0, 6]) & _lib.F_OVERFLOW:
            raise _lib.XqError("position exceeds the engine's capacity (>128 legal moves or >256 "
                               "pseudo-legal candidates)")
        return [unpack_move(m) for m in dev.moves[0, :n].cpu().tolist()]

    # -- chess_env.py:253-406 ----------------------------------------------------------------
    def make_move(self, move: Move):
        fr, fc, tr, tc = (int(x) for x in move)
        if not (0 <= fr < BOARD_SIZE and 0 <= tr < BOARD_SIZE and 0 <= fc < BOARD_WIDTH and 0 <= tc < BOARD_WIDTH):
            raise IndexError(f"move {move} is off the 10x9 board")  # numpy raises here too (>= size)
        dev = _Device1.get()
        self._upload(dev, with_hist=True)
        dev.move.fill_(pack_move(move))
        check(dev.lib.xq_step(_ptr(dev.board), _ptr(dev.meta), _ptr(dev.hist), dev.hist_cap,
                              _ptr(dev.move), _ptr(dev.reward), _ptr(dev.flags), None, None, 1,
                              _stream()))
        m = dev.meta.cpu().numpy().view(META_DTYPE).reshape(-1)[0]
        flags = int(dev.flags[0])
        reward = float(dev.reward[0])
        self.board = dev.board[0, :90].cpu().numpy().reshape(BOARD_SIZE, BOARD_WIDTH).copy()
        pos = lambda s: None if s < 0 else (int(s) // 9, int(s) % 9)
        self.red_king_pos, self.black_king_pos = pos(m["red_king"]), pos(m["black_king"])
        self.no_capture_count = int(m["no_capture"])
        self.consecutive_checks = int(m["consecutive_checks"])
        n_hist = int(m["hist_len"])
        if n_hist > len(self.position_history):
            self.position_history.append(int(np.uint64(dev.hist[0, n_hist - 1].item() & 0xFFFFFFFFFFFFFFFF)))
        self.check_history.append(bool(int(m["check_bits"]) & 1))
        self.chase_history.append([])
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
        self._upload(dev)
        check(dev.lib.xq_query_checks(_ptr(dev.board), _ptr(dev.meta), _ptr(dev.q), 1, _stream()))
        return dev.q[0].cpu().tolist()

    def _is_in_check(self, player) -> bool:
        q = self._query()
        return bool(q[0] if player == 1 else q[1])

    def _are_kings_facing(self) -> bool:
        return bool(self._query()[2])

    def _get_position_hash(self) -> int:
        dev = _Device1.get()
        self._upload(dev)
        check(dev.lib.xq_position_hash(_ptr(dev.board), _ptr(dev.meta), _ptr(dev.key), 1, _stream()))
        return int(dev.key[0].item()) & 0xFFFFFFFFFFFFFFFF

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

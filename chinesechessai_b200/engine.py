"""Batched Xiangqi boards resident in HBM, driven through the C ABI.

``BoardBatch`` owns struct-of-arrays game state as torch CUDA tensors (torch is
only the allocator / stream provider) and exposes the batched counterparts of
the reference's ``ChineseChess`` methods (chess_env.py): ``legal_moves`` (:76),
``step`` (``make_move`` :253) and the fused ``playout``.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (BOARD_STRIDE, MAX_MOVES, META_DTYPE, NSQ, PLAYOUT_RESULT_DTYPE, check)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def pack_move(move) -> int:
    fr, fc, tr, tc = move
    return (fr * 9 + fc) * 90 + tr * 9 + tc


def unpack_move(m: int) -> Tuple[int, int, int, int]:
    f, t = divmod(int(m), 90)
    return (f // 9, f % 9, t // 9, t % 9)


def decode_step_flags(flags: np.ndarray) -> Dict[str, np.ndarray]:
    f = flags.astype(np.int32)
    winner = ((f >> 2) & 3) - 1
    return dict(done=(f & 1).astype(bool), reward_is_int=((f >> 1) & 1).astype(bool),
                winner=np.where(winner == 2, _lib.WINNER_NONE, winner), reason=(f >> 4) & 15)


class BoardBatch:
    """``n`` independent games as SoA tensors on one GPU."""

    def __init__(self, n: int, device: Optional[torch.device] = None, hist_cap: int = 128):
        self.lib = _lib.load()
        _lib.require_device()
        if not torch.cuda.is_available():
            raise _lib.XqError("torch sees no CUDA device; the engine has no CPU fallback")
        self.n = int(n)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.hist_cap = int(hist_cap)
        d = self.device
        self.board = torch.zeros((self.n, BOARD_STRIDE), dtype=torch.int8, device=d)
        self.meta = torch.zeros((self.n, 32), dtype=torch.uint8, device=d)
        self.pos_hist = torch.zeros((self.n, self.hist_cap), dtype=torch.int64, device=d)
        self.moves = torch.zeros((self.n, MAX_MOVES), dtype=torch.int16, device=d)
        self.n_moves = torch.zeros((self.n,), dtype=torch.int16, device=d)
        self.reward = torch.zeros((self.n,), dtype=torch.float64, device=d)
        self.flags = torch.zeros((self.n,), dtype=torch.uint8, device=d)
        self.reset()

    # -- state ---------------------------------------------------------------
    def reset(self) -> None:
        with torch.cuda.device(self.device):
            check(self.lib.xq_reset(_ptr(self.board), _ptr(self.meta), self.n, _stream()))

    def set_state(self, boards: np.ndarray, meta: np.ndarray,
                  pos_hist: Optional[np.ndarray] = None) -> None:
        """Upload host state: boards int8[n,90|96] (or [n,10,9]), meta META_DTYPE[n]."""
        b = np.asarray(boards, dtype=np.int8).reshape(self.n, -1)
        full = np.zeros((self.n, BOARD_STRIDE), np.int8)
        full[:, :b.shape[1]] = b
        assert meta.dtype == META_DTYPE and meta.shape == (self.n,)
        self.board.copy_(torch.from_numpy(full))
        self.meta.copy_(torch.from_numpy(meta.view(np.uint8).reshape(self.n, 32)))
        if pos_hist is not None:
            h = np.zeros((self.n, self.hist_cap), np.uint64)
            h[:, :pos_hist.shape[1]] = pos_hist
            self.pos_hist.copy_(torch.from_numpy(h.view(np.int64)))

    def boards_host(self) -> np.ndarray:
        return self.board[:, :NSQ].cpu().numpy()

    def meta_host(self) -> np.ndarray:
        return self.meta.cpu().numpy().view(META_DTYPE).reshape(self.n)

    def pos_hist_host(self) -> np.ndarray:
        return self.pos_hist.cpu().numpy().view(np.uint64)

    def position_hash(self) -> torch.Tensor:
        out = torch.empty((self.n,), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.xq_position_hash(_ptr(self.board), _ptr(self.meta), _ptr(out), self.n,
                                            _stream()))
        return out

    # -- rules -----------------------------------------------------------------
    def legal_moves(self, in_check: Optional[torch.Tensor] = None):
        """get_legal_moves for every game -> (moves int16[n,128], n_moves int16[n])."""
        with torch.cuda.device(self.device):
            check(self.lib.xq_legal_moves(_ptr(self.board), _ptr(self.meta), _ptr(self.moves),
                                          _ptr(self.n_moves), _ptr(in_check), self.n, _stream()))
        return self.moves, self.n_moves

    def step(self, move: torch.Tensor, want_next: bool = False):
        """make_move for every game (move < 0 freezes a game) -> (reward f64[n], flags u8[n])."""
        assert move.dtype == torch.int16 and move.shape == (self.n,) and move.is_cuda
        with torch.cuda.device(self.device):
            check(self.lib.xq_step(_ptr(self.board), _ptr(self.meta), _ptr(self.pos_hist),
                                   self.hist_cap, _ptr(move), _ptr(self.reward), _ptr(self.flags),
                                   _ptr(self.moves) if want_next else None,
                                   _ptr(self.n_moves) if want_next else None, self.n, _stream()))
        return self.reward, self.flags

    def step_pick(self, seed: int, ply: int, first_game_id: int = 0, capture_bias: int = 0,
                  picked: Optional[torch.Tensor] = None):
        """One ply of the random-playout loop in ONE launch (xq_step_pick): pick from the legal
        list in ``self.moves`` (left by ``legal_moves()`` or the previous call), make_move, next
        legal list written back.  -> (reward f64[n], flags u8[n])."""
        with torch.cuda.device(self.device):
            check(self.lib.xq_step_pick(_ptr(self.board), _ptr(self.meta), _ptr(self.pos_hist),
                                        self.hist_cap, _ptr(self.moves), _ptr(self.n_moves), seed,
                                        first_game_id, ply, capture_bias, _ptr(self.reward),
                                        _ptr(self.flags), _ptr(picked), self.n, _stream()))
        return self.reward, self.flags

    def pick(self, seed: int, ply: int, first_game_id: int = 0, capture_bias: int = 0,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.n,), dtype=torch.int16, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.xq_pick_moves(_ptr(self.board), _ptr(self.meta), _ptr(self.moves),
                                         _ptr(self.n_moves), seed, first_game_id, ply, capture_bias,
                                         _ptr(out), self.n, _stream()))
        return out

    def playout(self, seed: int, max_plies: int = 70, first_game_id: int = 0,
                capture_bias: int = 0, trace: bool = False, results: Optional[torch.Tensor] = None):
        """Fused random playout (one launch).  Returns results tensor uint8[n,40] (view with
        PLAYOUT_RESULT_DTYPE on the host) and, if ``trace``, a dict of per-ply device tensors."""
        d = self.device
        if results is None:
            results = torch.zeros((self.n, 40), dtype=torch.uint8, device=d)
        tr = None
        ptrs = [None] * 6
        if trace:
            tr = dict(
                moves=torch.zeros((self.n, max_plies, MAX_MOVES), dtype=torch.int16, device=d),
                n=torch.zeros((self.n, max_plies), dtype=torch.int16, device=d),
                pick=torch.zeros((self.n, max_plies), dtype=torch.int16, device=d),
                reward=torch.zeros((self.n, max_plies), dtype=torch.float64, device=d),
                flags=torch.zeros((self.n, max_plies), dtype=torch.uint8, device=d),
                boards=torch.zeros((self.n, max_plies, NSQ), dtype=torch.int8, device=d))
            ptrs = [_ptr(tr[k]) for k in ("moves", "n", "pick", "reward", "flags", "boards")]
        with torch.cuda.device(d):
            check(self.lib.xq_playout(_ptr(self.board), _ptr(self.meta), _ptr(self.pos_hist),
                                      self.hist_cap, seed, first_game_id, max_plies, capture_bias,
                                      _ptr(results), *ptrs, self.n, _stream()))
        return (results, tr) if trace else results


def results_host(results: torch.Tensor) -> np.ndarray:
    return results.cpu().numpy().view(PLAYOUT_RESULT_DTYPE).reshape(-1)


def playout_host(boards: np.ndarray, meta: np.ndarray, seed: int, max_plies: int = 70,
                 first_game_id: int = 0, capture_bias: int = 0, device: int = 0,
                 results: Optional[np.ndarray] = None) -> np.ndarray:
    """End-to-end playout from HOST buffers through ``xq_playout_host`` (H2D + kernel + D2H).
    ``boards`` int8[n,96] and ``meta`` META_DTYPE[n] are updated in place."""
    lib = _lib.load()
    n = len(meta)
    assert boards.dtype == np.int8 and boards.shape == (n, BOARD_STRIDE) and boards.flags.c_contiguous
    assert meta.dtype == META_DTYPE and meta.flags.c_contiguous
    if results is None:
        results = np.zeros(n, PLAYOUT_RESULT_DTYPE)
    check(lib.xq_playout_host(boards.ctypes.data, meta.ctypes.data, seed, first_game_id, max_plies,
                              capture_bias, results.ctypes.data, n, device))
    return results


def encode_planes(board: torch.Tensor, player: torch.Tensor, out: Optional[torch.Tensor] = None,
                  dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """ChessNet.encode_board (neural_network.py:128-146) for a batch.
    board int8[n,>=90] (row stride in elements), player int8[n] (any stride)."""
    lib = _lib.load()
    n = board.shape[0]
    if out is None:
        out = torch.empty((n, 15, 10, 9), dtype=dtype, device=board.device)
    assert out.dtype in (torch.float32, torch.bfloat16) and out.is_contiguous()
    with torch.cuda.device(board.device):
        check(lib.xq_encode_planes(_ptr(board), board.stride(0), _ptr(player), player.stride(0),
                                   _ptr(out), 1 if out.dtype == torch.bfloat16 else 0, n, _stream()))
    return out


def encode_planes_nhwc16(board: torch.Tensor, player: torch.Tensor) -> torch.Tensor:
    """The same planes as a bf16 tensor of shape [n,16,10,9] in channels_last memory format
    (channel 15 is zero padding): what the folded inference network consumes, written by the
    kernel in that layout instead of NCHW + a layout-conversion pass."""
    lib = _lib.load()
    n = board.shape[0]
    out = torch.empty((n, 16, 10, 9), dtype=torch.bfloat16, device=board.device,
                      memory_format=torch.channels_last)
    with torch.cuda.device(board.device):
        check(lib.xq_encode_planes_nhwc16(_ptr(board), board.stride(0), _ptr(player), player.stride(0),
                                          _ptr(out), n, _stream()))
    return out


def stem_lookup(board: torch.Tensor, player: torch.Tensor, table: torch.Tensor, bias: torch.Tensor
                ) -> torch.Tensor:
    """encode_board + conv1/bn1/ReLU of ChessNet in one kernel (xq_stem_lookup_bf16): the planes are
    one-hot, so the first convolution is a sum of weight columns picked by the neighbouring
    pieces.  table bf16 [9,16,C], bias float32 [C] -> bf16 [n,C,10,9] channels-last."""
    lib = _lib.load()
    n, ch = board.shape[0], table.shape[2]
    assert table.dtype == torch.bfloat16 and table.shape[:2] == (9, 16) and table.is_contiguous()
    assert bias.dtype == torch.float32 and bias.shape == (2, 90, ch) and bias.is_contiguous()
    out = torch.empty((n, ch, 10, 9), dtype=torch.bfloat16, device=board.device,
                      memory_format=torch.channels_last)
    with torch.cuda.device(board.device):
        check(lib.xq_stem_lookup_bf16(_ptr(board), board.stride(0), _ptr(player), player.stride(0),
                                      _ptr(table), _ptr(bias), _ptr(out), ch, n, _stream()))
    return out


def policy_priors(logits: torch.Tensor, moves: torch.Tensor, n_moves: torch.Tensor,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ChessNet._logits_to_move_probs (neural_network.py:148-169) for a batch."""
    lib = _lib.load()
    n = logits.shape[0]
    assert logits.stride(1) == 1 and logits.shape[1] >= _lib.POLICY  # padded heads allowed
    assert logits.dtype in (torch.float32, torch.bfloat16)
    if out is None:
        out = torch.empty((n, MAX_MOVES), dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        check(lib.xq_policy_priors(_ptr(logits), 1 if logits.dtype == torch.bfloat16 else 0,
                                   logits.stride(0), _ptr(moves), moves.stride(0), _ptr(n_moves),
                                   _ptr(out), n, _stream()))
    return out


def bias_residual_relu(y: torch.Tensor, x: torch.Tensor, bias: torch.Tensor,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """relu(y + bias[c] + x) for bf16 channels-last activations (ResidualBlock epilogue,
    neural_network.py:181-187), one HBM pass in the CUDA kernel xq_bias_residual_relu_bf16."""
    lib = _lib.load()
    assert y.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and bias.dtype == torch.bfloat16
    assert y.shape == x.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert x.is_contiguous(memory_format=torch.channels_last) and bias.is_contiguous()
    if out is None:
        out = torch.empty_like(y)  # preserves channels_last
    with torch.cuda.device(y.device):
        check(lib.xq_bias_residual_relu_bf16(_ptr(y), _ptr(x), _ptr(bias), _ptr(out), y.numel(),
                                             y.shape[1], _stream()))
    return out

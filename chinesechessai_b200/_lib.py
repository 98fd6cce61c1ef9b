"""ctypes binding of libxq_b200.so — the C ABI declared in include/xq_b200.h.

There is no CPU fallback: if the CUDA library has not been built the import of
the engine fails loudly, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxq_b200.so")

NSQ = 90
BOARD_STRIDE = 96
MAX_MOVES = 128
WINNER_NONE = 2
PLANES = 15
POLICY = 8100
ABI_VERSION = 1

F_OVERFLOW = 1

REASON_NONE, REASON_KING_CAPTURE, REASON_CHECKMATE, REASON_REPETITION, REASON_FIFTY, \
    REASON_STALEMATE, REASON_PERPETUAL_CHECK, REASON_PERPETUAL_CHASE, REASON_MOVE_CAP = range(9)

# host mirror of xq_meta (32 bytes)
META_DTYPE = np.dtype([
    ("player", "i1"), ("winner", "i1"), ("reason", "u1"), ("done", "u1"),
    ("red_king", "i1"), ("black_king", "i1"), ("flags", "u1"), ("reserved", "u1"),
    ("move_count", "<i4"), ("no_capture", "<i4"), ("consecutive_checks", "<i4"),
    ("hist_len", "<i4"), ("check_bits", "<u4"), ("check_len", "<i4")])
assert META_DTYPE.itemsize == 32

# host mirror of xq_playout_result (40 bytes)
PLAYOUT_RESULT_DTYPE = np.dtype([
    ("plies", "<i4"), ("winner", "<i4"), ("reason", "<i4"), ("max_legal", "<i4"),
    ("reward_sum", "<f8"), ("digest", "<u8"), ("final_hash", "<u8")])
assert PLAYOUT_RESULT_DTYPE.itemsize == 40


class XqError(RuntimeError):
    pass


_vp, _i, _u32, _u64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64

# name -> (restype, argtypes); mirrors include/xq_b200.h one to one
SIGNATURES = {
    "xq_abi_version": (_i, []),
    "xq_last_error": (C.c_char_p, []),
    "xq_device_count": (_i, []),
    "xq_launch_count": (C.c_int64, []),
    "xq_reset": (_i, [_vp, _vp, _i, _vp]),
    "xq_position_hash": (_i, [_vp, _vp, _vp, _i, _vp]),
    "xq_legal_moves": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "xq_query_checks": (_i, [_vp, _vp, _vp, _i, _vp]),
    "xq_step": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "xq_step_pick": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _u64, _u32, _u32, _i, _vp, _vp, _vp, _i, _vp]),
    "xq_pick_moves": (_i, [_vp, _vp, _vp, _vp, _u64, _u32, _u32, _i, _vp, _i, _vp]),
    "xq_playout": (_i, [_vp, _vp, _vp, _i, _u64, _u32, _i, _i, _vp,
                        _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "xq_debug_playout_timing": (_i, [_vp]),
    "xq_playout_host": (_i, [_vp, _vp, _u64, _u32, _i, _i, _vp, _i, _i]),
    "xq_encode_planes": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _vp]),
    "xq_encode_planes_nhwc16": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp]),
    "xq_policy_priors": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "xq_bias_residual_relu_bf16": (_i, [_vp, _vp, _vp, _vp, C.c_int64, _i, _vp]),
    "xq_stem_lookup_bf16": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "xq_mcts_tree_bytes": (C.c_int64, [_i]),
    "xq_mcts_init": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp]),
    "xq_mcts_select": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "xq_mcts_backup": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "xq_mcts_backup_rows": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "xq_compact_leaves": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "xq_gather_leaves": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xq_mcts_root_visits": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp]),
    "xq_sample_moves": (_i, [_vp, _vp, _vp, C.c_double, _u64, _u32, _u32, _vp, _i, _vp]),
    "xq_selfplay_commit": (_i, [_vp] * 15 + [_i, _vp]),
    "xq_selfplay_finish": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "xq_hash_eval": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
}

_lib = None


def load():
    """Load libxq_b200.so (raises XqError if it was never built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XqError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C chinesechessai_b200/csrc` (needs nvcc, sm_100a). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    got = lib.xq_abi_version()
    if got != ABI_VERSION:
        raise XqError(f"ABI mismatch: library {got}, binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise XqError(load().xq_last_error().decode("utf-8", "replace") or f"xq error {rc}")


def require_device() -> int:
    n = load().xq_device_count()
    if n <= 0:
        raise XqError("no CUDA device visible: " + load().xq_last_error().decode())
    return n

"""CPU: the kernel's per-lane legality logic (xq_rules.cuh: attacked / suicide / gen_item,
compiled for the host by tests/host_mirror) against the goldens and fuzzed against the oracle.
This is a test harness for the device code, not a product path."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def hm():
    from tests.host_mirror.build import build
    lib = C.CDLL(build())
    lib.xqh_legal_moves.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.xqh_in_check.argtypes = [C.c_void_p] + [C.c_int] * 4
    lib.xqh_in_check_dirs.argtypes = [C.c_void_p] + [C.c_int] * 4
    lib.xqh_check_fast.argtypes = [C.c_void_p] + [C.c_int] * 3
    lib.xqh_gen_piece_mismatches.argtypes = [C.c_void_p, C.c_int]
    lib.xqh_king_move_mismatches.argtypes = [C.c_void_p] + [C.c_int] * 3
    lib.xqh_position_change.argtypes = [C.c_int] * 5
    lib.xqh_position_change.restype = C.c_double
    return lib


def _legal(hm, board, player, red, black):
    """Kernel Phase B (mode 1) must agree with the test-everything path (mode 0)."""
    b = np.ascontiguousarray(board, np.int8).reshape(90)
    out = []
    for mode in (0, 1, 2):
        mv = np.zeros(128, np.int16)
        nc = C.c_int(0)
        n = hm.xqh_legal_moves(b.ctypes.data, int(player), int(red), int(black), mv.ctypes.data,
                               C.byref(nc), mode)
        out.append(mv[:n].copy())
    assert np.array_equal(out[0], out[1]), "relevance filter changed the legal list"
    assert np.array_equal(out[0], out[2]), "regular-position fast path changed the legal list"
    return out[1], nc.value


def test_golden_positions(hm, golden):
    P = golden.positions
    off = P["legal_offset"]
    for i in range(len(P["player"])):
        mv, _ = _legal(hm, P["board"][i], P["player"][i], P["red"][i], P["black"][i])
        assert np.array_equal(mv, P["legal"][off[i]:off[i + 1]]), i
        b = np.ascontiguousarray(P["board"][i])
        pl = int(P["player"][i])
        assert hm.xqh_in_check(b.ctypes.data, pl, pl, int(P["red"][i]), int(P["black"][i])) == P["chk_self"][i], i
        assert hm.xqh_in_check(b.ctypes.data, -pl, pl, int(P["red"][i]), int(P["black"][i])) == P["chk_opp"][i], i
        for who in (pl, -pl):
            assert hm.xqh_in_check_dirs(b.ctypes.data, who, pl, int(P["red"][i]), int(P["black"][i])) == \
                hm.xqh_in_check(b.ctypes.data, who, pl, int(P["red"][i]), int(P["black"][i])), i


@pytest.mark.parametrize("bias", [0, 200])
def test_fuzz_playouts_vs_oracle(hm, xo, bias):
    max_cand = 0
    n_fast = [0]
    for g in range(150):
        e = xo.Env()
        for ply in range(70):
            lm = e.legal_moves_packed()
            mv, nc = _legal(hm, e.board, e.s.player, e.s.red_king, e.s.black_king)
            max_cand = max(max_cand, nc)
            assert np.array_equal(mv, lm), (g, ply)
            assert nc == e.pseudo_count()
            b = np.ascontiguousarray(e.board.reshape(90))
            for who in (1, -1):
                assert hm.xqh_in_check(b.ctypes.data, who, e.s.player, e.s.red_king, e.s.black_king) == \
                    int(e.is_in_check(who)), (g, ply, who)
                assert hm.xqh_in_check_dirs(b.ctypes.data, who, e.s.player, e.s.red_king, e.s.black_king) == \
                    int(e.is_in_check(who)), (g, ply, who)
            # make_move's check test (:317) = the side to move's king under the PREVIOUS mover's
            # geometry: the mask form used on regular positions against the general probes
            fast = hm.xqh_check_fast(b.ctypes.data, e.s.player, e.s.red_king, e.s.black_king)
            assert fast == hm.xqh_in_check(b.ctypes.data, e.s.player, -e.s.player, e.s.red_king, e.s.black_king), (g, ply)
            n_fast[0] += fast >= 0
            assert hm.xqh_gen_piece_mismatches(b.ctypes.data, e.s.player) == 0, (g, ply)
            # the king's own moves: table test == the general probes (regular positions only)
            assert hm.xqh_king_move_mismatches(b.ctypes.data, e.s.player, e.s.red_king, e.s.black_king) == 0, (g, ply)
            if len(lm) == 0:
                break
            idx = xo.lib().xqo_pick_move(e.s, lm.ctypes.data, len(lm), 99, g, ply, bias)
            _, _, done = e.make_move(int(lm[idx]))
            if done:
                break
    assert max_cand <= 128 and n_fast[0] > 5000       # play never leaves the regular path


def test_fuzz_arbitrary_boards_vs_oracle(hm, xo):
    rng = np.random.default_rng(123)
    n_regular = 0
    for it in range(4000):
        board = np.zeros(90, np.int8)
        k = int(rng.integers(2, 30))
        sq = rng.choice(90, size=k, replace=False)
        board[sq] = rng.integers(1, 8, size=k) * rng.choice([-1, 1], size=k)
        player = int(rng.choice([-1, 1]))
        # the per-piece generator of the lane-pair engine == the per-slot one, also for piece codes
        # outside 1..7 (nothing is generated for them)
        odd = board.copy()
        if it % 8 == 0:
            odd[sq[0]] = int(rng.integers(8, 100)) * int(rng.choice([-1, 1]))
        for who in (1, -1):
            assert hm.xqh_gen_piece_mismatches(odd.ctypes.data, who) == 0, it

        def cache(code):
            u = rng.random()
            w = np.flatnonzero(board == code)
            if u < 0.6 and len(w):
                return int(w[0])
            if u < 0.8:
                return int(rng.integers(0, 90))
            return -1
        red, black = cache(1), cache(-1)
        pos = lambda s: None if s < 0 else (s // 9, s % 9)
        e = xo.Env().load(board.reshape(10, 9), player, 0, None, pos(red), pos(black))
        if e.pseudo_count() > 256:
            continue
        mv, _ = _legal(hm, board, player, red, black)
        assert np.array_equal(mv, e.legal_moves_packed()), it
        fast = hm.xqh_check_fast(board.ctypes.data, player, red, black)
        if fast >= 0:
            assert fast == hm.xqh_in_check(board.ctypes.data, player, -player, red, black), it
        km = hm.xqh_king_move_mismatches(board.ctypes.data, player, red, black)
        assert km <= 0 and (km < 0) == (fast < 0), it
        n_regular += km == 0
    assert n_regular > 150


def test_fuzz_king_moves_on_regular_boards(hm):
    """king_move_fast() == suicide() for the king's own moves on dense regular positions: both
    kings in their palaces (cached), no other K/A/B of the enemy, rooks / cannons / knights / pawns
    of both sides scattered — around the king as well, so that screens, legs, adjacent pawns from
    every side and the kings' file are all exercised."""
    rng = np.random.default_rng(77)
    n_cand = 0
    for it in range(6000):
        board = np.zeros(90, np.int8)
        player = int(rng.choice([-1, 1]))
        rk = int(rng.integers(7, 10)) * 9 + int(rng.integers(3, 6))
        bk = int(rng.integers(0, 3)) * 9 + int(rng.integers(3, 6))
        board[rk], board[bk] = 1, -1
        own = rk if player == 1 else bk
        kr, kc = divmod(own, 9)
        k = int(rng.integers(3, 26))
        near = [r * 9 + c for r in range(max(0, kr - 3), min(10, kr + 4)) for c in range(max(0, kc - 3), min(9, kc + 4))]
        for _ in range(k):
            s = int(rng.choice(near)) if rng.random() < 0.6 else int(rng.integers(0, 90))
            if board[s] == 0:
                board[s] = int(rng.choice([4, 5, 6, 7])) * int(rng.choice([-1, 1]))
        if rng.random() < 0.3:   # own advisors / bishops are harmless to the regular test
            s = int(rng.choice(near))
            if board[s] == 0:
                board[s] = player * int(rng.choice([2, 3]))
        km = hm.xqh_king_move_mismatches(board.ctypes.data, player, rk, bk)
        assert km == 0, (it, board.reshape(10, 9), player)
        n_cand += 1
    assert n_cand == 6000

#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED Python reference.

Runs only in the authoring container (needs /root/reference; it does not exist
on the GPU box).  Every ply produced here is ALSO stepped on the C oracle
(oracle/xq_oracle.c) and compared field by field, so this script is both the
fixture generator and the differential fuzzer that pins the oracle.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden.py [--quick]

Outputs (committed): tests/golden/playouts.npz, positions.npz, mcts.npz,
kats.json, MANIFEST.json.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("XQ_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

M64 = (1 << 64) - 1
SEED = 0x5EED


# ---- independent Python restatement of the shared pick rule / digest ---------
def mix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def pos_hash(board: np.ndarray, player: int) -> int:
    h = mix64(0x7000 + (0 if player == 1 else 1))
    flat = board.reshape(90)
    for sq in range(90):
        p = int(flat[sq])
        if p:
            h ^= mix64((p + 8) * 128 + sq)
    return h


def philox4x32(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, \
            ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def pick_index(board, packed_moves, seed, game_id, ply, bias):
    x = philox4x32(game_id, ply, 0, 0, seed & 0xFFFFFFFF, seed >> 32)
    n = len(packed_moves)
    if bias > 0 and (x[1] & 0xFF) < bias:
        caps = [i for i, m in enumerate(packed_moves) if board.reshape(90)[m % 90] != 0]
        if caps:
            return caps[x[0] % len(caps)]
    return x[0] % n


def pack(mv):
    return (mv[0] * 9 + mv[1]) * 90 + mv[2] * 9 + mv[3]


def reason_code(s):
    if s is None:
        return 0
    if s.endswith("吃掉对方将帅"):
        return 1
    if s.startswith("将死"):
        return 2
    if s == "三次重复局面判和":
        return 3
    if s == "50回合无吃子判和":
        return 4
    if s.startswith("困毙"):
        return 5
    if s.startswith("长将判负"):
        return 6
    if s.startswith("长捉判负"):
        return 7
    if s.startswith("超过"):
        return 8
    raise ValueError(s)


def dbits(x: float) -> int:
    return int(np.float64(x).view(np.uint64))


def ply_digest(ply, packed, pick_move, board_after, player_after, reward, done, winner, reason,
               is_int):
    lsum = 0
    for i, m in enumerate(packed):
        lsum = (lsum + (m + 1) * (2 * i + 1)) & 0xFFFFFFFF
    a = lsum | (len(packed) << 32) | (pick_move << 40) | ((ply + 1) << 54)
    wn = 2 if winner is None else winner
    c = (1 if done else 0) | ((wn + 2) << 8) | (reason << 16) | ((1 if is_int else 0) << 24)
    return (a * 0x9E3779B97F4A7C15 + dbits(reward) * 0xC2B2AE3D27D4EB4F + c * 0x165667B19E3779F9 +
            pos_hash(board_after, player_after) * 0x27D4EB2F165667C5) & M64


def import_reference():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import chess_env  # noqa
        import self_play  # noqa
    return chess_env, self_play


def sq(pos):
    return -1 if pos is None else pos[0] * 9 + pos[1]


def compare_env(ref, orc, ctx):
    """Field-by-field differential check reference env vs oracle env."""
    assert np.array_equal(ref.board, orc.board), ("board", ctx)
    assert ref.current_player == orc.current_player, ("player", ctx)
    assert ref.move_count == orc.s.move_count, ("move_count", ctx)
    assert ref.winner == orc.winner, ("winner", ctx, ref.winner, orc.winner)
    assert reason_code(ref.end_reason) == orc.s.reason, ("reason", ctx)
    assert sq(ref.red_king_pos) == orc.s.red_king, ("red_king", ctx)
    assert sq(ref.black_king_pos) == orc.s.black_king, ("black_king", ctx)
    assert ref.no_capture_count == orc.s.no_capture, ("no_capture", ctx)
    assert ref.consecutive_checks == orc.s.consecutive_checks, ("cc", ctx)
    assert [bool(x) for x in ref.check_history] == orc.check_history, ("check_hist", ctx)
    assert len(ref.position_history) == orc.s.pos_len, ("pos_len", ctx)
    # hash equality structure (reference hashes are salted: compare partitions)
    rh, oh = ref.position_history, orc.position_history
    if rh:
        assert [rh.index(x) for x in rh] == [oh.index(x) for x in oh], ("pos_hist", ctx)


# ---- playout traces ----------------------------------------------------------------
def play_one(args):
    game_id, bias, full = args
    chess_env, _ = import_reference()
    from oracle import xq_oracle as xo
    ref = chess_env.ChineseChess()
    orc = xo.Env()
    rec = dict(game_id=game_id, bias=bias, moves=[], n=[], pick=[], reward=[], flags=[],
               boards=[], kings=[], digest=0, reward_sum=0.0, max_pseudo=0)
    ply = 0
    while ply < 70:
        legal = ref.get_legal_moves()
        packed = [pack(m) for m in legal]
        o_packed = [int(x) for x in orc.legal_moves_packed()]
        assert packed == o_packed, ("legal", game_id, ply)
        rec["max_pseudo"] = max(rec["max_pseudo"], orc.pseudo_count())
        if not legal:
            break
        idx = pick_index(ref.board, packed, SEED, game_id, ply, bias)
        o_idx = xo.lib().xqo_pick_move(orc.s, np.array(packed, np.int16).ctypes.data, len(packed),
                                       SEED, game_id, ply, bias)
        assert idx == o_idx, ("pick", game_id, ply)
        (_b, _p), reward, done = ref.make_move(legal[idx])
        o_reward, o_is_int, o_done = orc.make_move(packed[idx])
        is_int = isinstance(reward, int)
        assert repr(float(reward)) == repr(float(o_reward)), ("reward", game_id, ply, reward, o_reward)
        assert is_int == o_is_int, ("is_int", game_id, ply, reward)
        assert bool(done) == o_done, ("done", game_id, ply)
        compare_env(ref, orc, (game_id, ply))
        rc = reason_code(ref.end_reason)
        wn = 2 if ref.winner is None else ref.winner
        rec["reward_sum"] += float(reward)
        rec["digest"] = mix64(rec["digest"] ^ ply_digest(
            ply, packed, packed[idx], ref.board, ref.current_player, float(reward), done,
            ref.winner, rc, is_int))
        if full:
            rec["moves"].append(packed)
            rec["boards"].append(ref.board.reshape(90).copy())
            rec["kings"].append((sq(ref.red_king_pos), sq(ref.black_king_pos),
                                 ref.consecutive_checks, ref.no_capture_count))
        rec["n"].append(len(packed))
        rec["pick"].append(packed[idx])
        rec["reward"].append(float(reward))
        rec["flags"].append((1 if done else 0) | ((1 if is_int else 0) << 1) | ((wn + 1) << 2) |
                            (rc << 4))
        ply += 1
        if done:
            break
    rec["plies"] = ply
    rec["winner"] = 2 if ref.winner is None else ref.winner
    rec["reason"] = reason_code(ref.end_reason)
    rec["end_reason"] = ref.end_reason
    rec["final_hash"] = pos_hash(ref.board, ref.current_player)
    return rec


def gen_playouts(pool, n_uniform, n_biased, n_full):
    jobs = [(g, 0, g < n_full) for g in range(n_uniform)]
    jobs += [(100000 + g, 192, g < n_full) for g in range(n_biased)]
    recs = pool.map(play_one, jobs, chunksize=1)
    summ = np.zeros(len(recs), dtype=[("game_id", "<u4"), ("bias", "<i4"), ("plies", "<i4"),
                                      ("winner", "<i4"), ("reason", "<i4"), ("max_legal", "<i4"),
                                      ("max_pseudo", "<i4"), ("reward_sum", "<f8"),
                                      ("digest", "<u8"), ("final_hash", "<u8")])
    for i, r in enumerate(recs):
        summ[i] = (r["game_id"], r["bias"], r["plies"], r["winner"], r["reason"],
                   max(r["n"]) if r["n"] else 0, r["max_pseudo"], r["reward_sum"], r["digest"],
                   r["final_hash"])
    full = [r for r in recs if r["moves"]]
    out = dict(summary=summ)
    # ragged full traces
    ply_off = np.cumsum([0] + [r["plies"] for r in full]).astype(np.int64)
    out["full_game_index"] = np.array([recs.index(r) for r in full], np.int32)
    out["full_ply_offset"] = ply_off
    out["full_n"] = np.concatenate([np.array(r["n"], np.int16) for r in full])
    out["full_pick"] = np.concatenate([np.array(r["pick"], np.int16) for r in full])
    out["full_reward"] = np.concatenate([np.array(r["reward"], np.float64) for r in full])
    out["full_flags"] = np.concatenate([np.array(r["flags"], np.uint8) for r in full])
    out["full_boards"] = np.concatenate([np.stack(r["boards"]) for r in full]).astype(np.int8)
    out["full_kings"] = np.concatenate([np.array(r["kings"], np.int16) for r in full])
    mv = [np.array(m, np.int16) for r in full for m in r["moves"]]
    out["full_moves"] = np.concatenate(mv)
    out["full_move_offset"] = np.cumsum([0] + [len(m) for m in mv]).astype(np.int64)
    reasons = {str(recs.index(r)): r["end_reason"] for r in full}
    return out, reasons, recs


# ---- arbitrary poked positions -------------------------------------------------------
def random_position(rng, kind):
    board = np.zeros((10, 9), np.int8)
    if kind == 0:  # anything anywhere (kings possibly outside the palace / missing / doubled)
        k = int(rng.integers(2, 24))
        sqs = rng.choice(90, size=k, replace=False)
        for s in sqs:
            code = int(rng.integers(1, 8)) * (1 if rng.random() < 0.5 else -1)
            board[s // 9, s % 9] = code
    else:  # plausible sparse endgame: kings in palaces, A/B on their own squares
        rk = (int(rng.integers(7, 10)), int(rng.integers(3, 6)))
        bk = (int(rng.integers(0, 3)), int(rng.integers(3, 6)))
        board[rk], board[bk] = 1, -1
        k = int(rng.integers(1, 14))
        for _ in range(k):
            s = int(rng.integers(0, 90))
            if board[s // 9, s % 9] == 0:
                t = int(rng.choice([4, 5, 6, 7, 7, 2, 3]))
                board[s // 9, s % 9] = t * (1 if rng.random() < 0.5 else -1)
    player = 1 if rng.random() < 0.5 else -1

    def find(code):
        w = np.argwhere(board == code)
        return None if len(w) == 0 else (int(w[0][0]), int(w[0][1]))
    red, black = find(1), find(-1)
    u = rng.random()
    if u < 0.15:  # stale / arbitrary caches
        red = (int(rng.integers(0, 10)), int(rng.integers(0, 9)))
    elif u < 0.25:
        red = None
    u = rng.random()
    if u < 0.15:
        black = (int(rng.integers(0, 10)), int(rng.integers(0, 9)))
    elif u < 0.25:
        black = None
    move_count = int(rng.choice([0, 5, 33, 68, 69, 70, 120]))
    no_cap = int(rng.choice([0, 3, 98, 99, 100]))
    cc = int(rng.integers(0, 5))
    nck = int(rng.choice([0, 5, 11, 12, 20]))
    ck = [bool(rng.random() < 0.85) for _ in range(nck)]
    return board, player, red, black, move_count, no_cap, cc, ck


def position_job(args):
    seed, count = args
    chess_env, _ = import_reference()
    from oracle import xq_oracle as xo
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(count):
        board, player, red, black, mc, ncap, cc, ck = random_position(rng, int(rng.integers(0, 2)))
        ref = chess_env.ChineseChess()
        ref.board = board.copy()
        ref.current_player = player
        ref.red_king_pos, ref.black_king_pos = red, black
        ref.move_count, ref.no_capture_count, ref.consecutive_checks = mc, ncap, cc
        ref.check_history = list(ck)
        orc = xo.Env().load(board, player, mc, None, red, black, ncap, cc, ck)
        legal = ref.get_legal_moves()
        packed = [pack(m) for m in legal]
        assert packed == [int(x) for x in orc.legal_moves_packed()], ("legal", seed, i)
        chk_self, chk_opp = bool(ref._is_in_check(player)), bool(ref._is_in_check(-player))
        facing = bool(ref._are_kings_facing())
        assert chk_self == orc.is_in_check(player) and chk_opp == orc.is_in_check(-player)
        assert facing == orc.kings_facing()
        row = dict(board=board.reshape(90).copy(), player=player, red=sq(red), black=sq(black),
                   mc=mc, ncap=ncap, cc=cc, ck=ck, legal=packed, chk_self=chk_self,
                   chk_opp=chk_opp, facing=facing, move=-1)
        if legal:
            j = int(rng.integers(0, len(legal)))
            (_b, _p), reward, done = ref.make_move(legal[j])
            o_reward, o_is_int, o_done = orc.make_move(packed[j])
            assert repr(float(reward)) == repr(float(o_reward)), ("reward", seed, i)
            assert isinstance(reward, int) == o_is_int and bool(done) == o_done
            compare_env(ref, orc, (seed, i))
            wn = 2 if ref.winner is None else ref.winner
            row.update(move=packed[j], reward=float(reward), is_int=isinstance(reward, int),
                       done=bool(done), winner=wn, reason=reason_code(ref.end_reason),
                       end_reason=ref.end_reason, board_after=ref.board.reshape(90).copy(),
                       red_after=sq(ref.red_king_pos), black_after=sq(ref.black_king_pos),
                       cc_after=ref.consecutive_checks, ncap_after=ref.no_capture_count,
                       check_flag=bool(ref.check_history[-1]))
        rows.append(row)
    return rows


def gen_positions(pool, n_jobs, per_job):
    rows = [r for chunk in pool.map(position_job, [(7000 + j, per_job) for j in range(n_jobs)],
                                    chunksize=1) for r in chunk]
    n = len(rows)
    out = dict(
        board=np.stack([r["board"] for r in rows]).astype(np.int8),
        player=np.array([r["player"] for r in rows], np.int8),
        red=np.array([r["red"] for r in rows], np.int16),
        black=np.array([r["black"] for r in rows], np.int16),
        mc=np.array([r["mc"] for r in rows], np.int32),
        ncap=np.array([r["ncap"] for r in rows], np.int32),
        cc=np.array([r["cc"] for r in rows], np.int32),
        ck_len=np.array([len(r["ck"]) for r in rows], np.int32),
        ck_bits=np.array([sum((1 << i) for i, v in enumerate(r["ck"][::-1]) if v) for r in rows],
                         np.uint32),  # bit 0 = most recent ply
        chk_self=np.array([r["chk_self"] for r in rows], np.uint8),
        chk_opp=np.array([r["chk_opp"] for r in rows], np.uint8),
        facing=np.array([r["facing"] for r in rows], np.uint8),
        move=np.array([r["move"] for r in rows], np.int16),
        reward=np.array([r.get("reward", 0.0) for r in rows], np.float64),
        is_int=np.array([r.get("is_int", False) for r in rows], np.uint8),
        done=np.array([r.get("done", False) for r in rows], np.uint8),
        winner=np.array([r.get("winner", 2) for r in rows], np.int8),
        reason=np.array([r.get("reason", 0) for r in rows], np.uint8),
        board_after=np.stack([r.get("board_after", r["board"]) for r in rows]).astype(np.int8),
        red_after=np.array([r.get("red_after", -1) for r in rows], np.int16),
        black_after=np.array([r.get("black_after", -1) for r in rows], np.int16),
        cc_after=np.array([r.get("cc_after", 0) for r in rows], np.int32),
        ncap_after=np.array([r.get("ncap_after", 0) for r in rows], np.int32),
        check_flag=np.array([r.get("check_flag", False) for r in rows], np.uint8),
    )
    lm = [np.array(r["legal"], np.int16) for r in rows]
    out["legal"] = np.concatenate(lm) if n else np.zeros(0, np.int16)
    out["legal_offset"] = np.cumsum([0] + [len(m) for m in lm]).astype(np.int64)
    return out




# ---- chase_history (dead bookkeeping of the reference, kept observable) ---------------------
def chase_job(args):
    """chess_env.py:262,344-345: the threat lists make_move appends to chase_history, for one
    random game (shared pick rule)."""
    game_id, bias = args
    chess_env, _ = import_reference()
    env = chess_env.ChineseChess()
    moves = []
    for ply in range(70):
        legal = env.get_legal_moves()
        if not legal:
            break
        packed = [pack(m) for m in legal]
        idx = pick_index(env.board, packed, SEED, game_id, ply, bias)
        moves.append(packed[idx])
        _, _, done = env.make_move(legal[idx])
        if done:
            break
    chase = [[[a[0] * 9 + a[1], b[0] * 9 + b[1]] for a, b in entry] for entry in env.chase_history]
    return dict(game_id=game_id, bias=bias, moves=moves, chase=chase)


def gen_chase(pool, quick):
    jobs = [(70000 + i, b) for i, b in enumerate([0, 160] if quick else [0, 0, 128, 192, 224, 255])]
    return pool.map(chase_job, jobs, chunksize=1)

# ---- compare_models.play_match --------------------------------------------------------------
def play_match_job(args):
    """The reference's UNCHANGED compare_models.play_match (compare_models.py:13-92) with two
    injected evaluators; the simulation count is the module default MCTS(network) picks up
    (self_play.MCTS_SIMULATIONS, set to `n_sims` here — a configuration value, not code).  Every
    make_move is logged through a wrapper so that the goldens hold the move sequences too."""
    seed, n_games, n_sims = args
    chess_env, self_play = import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        import compare_models
    self_play.MCTS_SIMULATIONS = n_sims
    log = []
    orig = chess_env.ChineseChess.make_move

    def logged(self, move):
        # only the game's own moves: the reference's MCTS also calls make_move, on env copies
        if sys._getframe(1).f_code.co_name == "play_match":
            log.append(pack(move))
        return orig(self, move)
    chess_env.ChineseChess.make_move = logged
    try:
        np.random.seed(seed)
        res = compare_models.play_match(StubNet(False), StubNet(True), num_games=n_games, verbose=False)
    finally:
        chess_env.ChineseChess.make_move = orig
    return dict(seed=seed, n_games=n_games, n_sims=n_sims, result=res, moves=log)


def gen_play_match(pool, quick):
    jobs = [(11, 2, 15)] if quick else [(11, 3, 15), (12, 2, 20)]
    return pool.map(play_match_job, jobs, chunksize=1)

# ---- MCTS ---------------------------------------------------------------------------
class StubNet:
    """Deterministic evaluator injected into the reference MCTS: priors are
    numpy.float32 (as _logits_to_move_probs returns), values are Python floats (.item())."""

    def __init__(self, flat):
        self.flat = flat
        self.calls = 0

    def predict_batch(self, items):
        from oracle import xq_oracle as xo
        self.calls += 1
        n = len(items)
        boards = np.stack([np.asarray(b, np.int8).reshape(90) for b, _, _ in items])
        players = np.array([p for _, p, _ in items], np.int32)
        moves = np.zeros((n, 128), np.int16)
        nm = np.zeros(n, np.int32)
        for i, (_, _, lm) in enumerate(items):
            nm[i] = len(lm)
            moves[i, :len(lm)] = [pack(m) for m in lm]
        pri, val = xo.hash_eval(boards, players, moves, nm, flat=self.flat)
        return [({m: pri[i, j] for j, m in enumerate(lm)}, float(val[i]))
                for i, (_, _, lm) in enumerate(items)]


def mcts_job(args):
    game_id, bias, root_ply, n_sims, flat = args
    chess_env, self_play = import_reference()
    from oracle import xq_oracle as xo
    ref = chess_env.ChineseChess()
    orc = xo.Env()
    for ply in range(root_ply):
        legal = ref.get_legal_moves()
        if not legal or ref.winner is not None:
            break
        packed = [pack(m) for m in legal]
        idx = pick_index(ref.board, packed, SEED, game_id, ply, bias)
        _, _, done = ref.make_move(legal[idx])
        orc.make_move(packed[idx])
        if done:
            break
    board0 = ref.board.copy()
    net = StubNet(flat)
    visits = self_play.MCTS(net, n_sims).search(ref, n_sims)
    assert np.array_equal(board0, ref.board)
    r_moves = [pack(m) for m in visits.keys()]
    r_vis = list(visits.values())
    o_moves, o_vis, st = xo.mcts_search(orc, n_sims, flat=flat)
    assert r_moves == [int(x) for x in o_moves], ("mcts moves", args)
    assert r_vis == [int(x) for x in o_vis], ("mcts visits", args, r_vis, list(o_vis))
    assert net.calls == st[2], ("predict_batch calls", args)
    wn = 2 if ref.winner is None else ref.winner
    return dict(board=ref.board.reshape(90).copy(), player=ref.current_player,
                mc=ref.move_count, winner=wn, red=sq(ref.red_king_pos),
                black=sq(ref.black_king_pos), ncap=ref.no_capture_count, n_sims=n_sims,
                flat=flat, moves=r_moves, visits=r_vis, stats=[int(x) for x in st])


def gen_mcts(pool, quick):
    jobs = []
    plies = [0, 7, 20, 41, 62, 66, 68, 69]
    sims = [15, 50] if quick else [15, 30, 50, 97]
    for i, rp in enumerate(plies):
        for n in sims:
            jobs.append((200 + i, 0, rp, n, False))
            jobs.append((100300 + i, 192, rp, n, False))
        jobs.append((200 + i, 0, rp, 50, True))
    if not quick:
        jobs.append((100300, 192, 30, 150, False))
        jobs.append((9, 0, 12, 8, False))   # n <= 8: root only, all-zero counts (B.4)
        jobs.append((9, 0, 12, 9, False))
    rows = pool.map(mcts_job, jobs, chunksize=1)
    out = dict(
        board=np.stack([r["board"] for r in rows]).astype(np.int8),
        player=np.array([r["player"] for r in rows], np.int8),
        mc=np.array([r["mc"] for r in rows], np.int32),
        winner=np.array([r["winner"] for r in rows], np.int8),
        red=np.array([r["red"] for r in rows], np.int16),
        black=np.array([r["black"] for r in rows], np.int16),
        ncap=np.array([r["ncap"] for r in rows], np.int32),
        n_sims=np.array([r["n_sims"] for r in rows], np.int32),
        flat=np.array([r["flat"] for r in rows], np.uint8),
        stats=np.array([r["stats"] for r in rows], np.int64),
    )
    mv = [np.array(r["moves"], np.int16) for r in rows]
    out["moves"] = np.concatenate(mv)
    out["visits"] = np.concatenate([np.array(r["visits"], np.int32) for r in rows])
    out["offset"] = np.cumsum([0] + [len(m) for m in mv]).astype(np.int64)
    return out


# ---- self_play_game (self_play.py:178-312) with the injected evaluator ----------------------
def selfplay_job(args):
    seed, n_sims, temperature, opponent = args
    chess_env, self_play = import_reference()
    np.random.seed(seed)
    net = StubNet(False)
    opp = StubNet(True) if opponent else None
    data, winner, reason = self_play.self_play_game(net, temperature=temperature,
                                                    num_simulations=n_sims, opponent_network=opp)
    return dict(seed=seed, n_sims=n_sims, temperature=temperature, opponent=bool(opponent),
                winner=int(winner), end_reason=reason,
                boards=[b.reshape(90).tolist() for b, _, _ in data],
                moves=[[pack(m) for m in d.keys()] for _, d, _ in data],
                probs=[[float(p) for p in d.values()] for _, d, _ in data],
                rewards=[float(r) for _, _, r in data])


def gen_selfplay(pool, quick):
    jobs = [(1, 15, 1.0, False), (2, 30, 1.0, False), (3, 30, 0.5, False), (4, 50, 1.0, False),
            (5, 30, 1.0, True), (6, 30, 0.001, False)]
    if quick:
        jobs = jobs[:2]
    return pool.map(selfplay_job, jobs, chunksize=1)


# ---- known-answer vectors ---------------------------------------------
def gen_kats():
    chess_env, _ = import_reference()
    from oracle import xq_oracle as xo
    kats = {}

    def run_line(name, moves, setup=None):
        ref = chess_env.ChineseChess()
        orc = xo.Env()
        if setup:
            setup(ref)
            orc.load(ref.board, ref.current_player, ref.move_count, ref.winner, ref.red_king_pos,
                     ref.black_king_pos, ref.no_capture_count, ref.consecutive_checks,
                     ref.check_history)
        start = dict(board=ref.board.reshape(90).tolist(), player=ref.current_player,
                     red=sq(ref.red_king_pos), black=sq(ref.black_king_pos),
                     check_history=[bool(x) for x in ref.check_history])
        rewards, dones, ints = [], [], []
        for mv in moves:
            (_b, _p), rw, dn = ref.make_move(tuple(mv))
            orw, oint, odn = orc.make_move(tuple(mv))
            assert repr(float(rw)) == repr(float(orw)) and bool(dn) == odn
            assert isinstance(rw, int) == oint
            compare_env(ref, orc, (name, mv))
            rewards.append(float(rw))
            ints.append(isinstance(rw, int))
            dones.append(bool(dn))
            if dn:
                break
        kats[name] = dict(start=start, moves=[list(m) for m in moves], rewards=rewards,
                          reward_is_int=ints, dones=dones, plies=len(rewards),
                          winner=ref.winner, end_reason=ref.end_reason,
                          reason=reason_code(ref.end_reason),
                          final_board=ref.board.reshape(90).tolist(),
                          consecutive_checks=ref.consecutive_checks,
                          last_check=bool(ref.check_history[-1]),
                          distinct_hashes=len(set(ref.position_history)))

    ref = chess_env.ChineseChess()
    kats["initial_legal_moves"] = [pack(m) for m in ref.get_legal_moves()]
    run_line("double_cannon_mate", [(7, 7, 7, 4), (0, 1, 2, 0), (7, 4, 3, 4), (2, 7, 7, 7),
                                    (7, 1, 5, 1), (2, 1, 2, 7), (5, 1, 5, 4)])
    run_line("knight_shuffle", [(9, 1, 7, 2), (0, 1, 2, 2), (7, 2, 9, 1), (2, 2, 0, 1)] * 18)
    run_line("quiet_knight", [(9, 1, 7, 2)])
    run_line("pawn_line", [(6, 4, 5, 4), (3, 4, 4, 4), (5, 4, 4, 4)])

    def cannon_takes_king(e):
        e.board[:] = 0
        e.board[0, 4], e.board[9, 4], e.board[0, 1], e.board[0, 2] = -1, 1, 6, -7
        e.black_king_pos, e.red_king_pos = (0, 4), (9, 4)
    run_line("cannon_takes_king", [(0, 1, 0, 4)], cannon_takes_king)

    def perpetual(e):
        e.board[:] = 0
        e.board[9, 4], e.board[0, 3], e.board[5, 0], e.board[2, 8] = 1, -1, 5, -5
        e.red_king_pos, e.black_king_pos = (9, 4), (0, 3)
        e.check_history = [True] * 11
    run_line("perpetual_check", [(5, 0, 5, 3)], perpetual)

    def repetition(e):
        e.board[:] = 0
        e.board[9, 4], e.board[9, 0], e.board[0, 3], e.board[0, 8] = 1, 5, -1, -5
        e.red_king_pos, e.black_king_pos = (9, 4), (0, 3)
    cyc = [(9, 0, 8, 0), (0, 8, 1, 8), (8, 0, 7, 0), (1, 8, 0, 8), (7, 0, 9, 0), (0, 8, 1, 8),
           (9, 0, 8, 0), (1, 8, 2, 8), (8, 0, 9, 0), (2, 8, 0, 8)]
    run_line("odd_cycle_repetition", cyc * 4, repetition)

    def fifty(e):
        e.board[:] = 0
        e.board[9, 4], e.board[9, 0], e.board[0, 3], e.board[0, 8] = 1, 5, -1, -5
        e.red_king_pos, e.black_king_pos = (9, 4), (0, 3)
        e.no_capture_count = 98
        e.move_count = 10
    run_line("fifty_move", [(9, 0, 8, 0), (0, 8, 1, 8), (8, 0, 7, 0)], fifty)

    # stale-cache / pawn-perspective single positions (A.3, A.4)
    def legal_of(setup):
        e = chess_env.ChineseChess()
        setup(e)
        o = xo.Env().load(e.board, e.current_player, 0, None, e.red_king_pos, e.black_king_pos)
        lm = [pack(m) for m in e.get_legal_moves()]
        assert lm == [int(x) for x in o.legal_moves_packed()]
        return dict(board=e.board.reshape(90).tolist(), player=e.current_player,
                    red=sq(e.red_king_pos), black=sq(e.black_king_pos), legal=lm,
                    in_check=bool(e._is_in_check(e.current_player)))

    def stale_a(e):
        e.board[:] = 0
        e.board[9, 4], e.board[0, 4], e.board[1, 4] = 1, -1, 7
        e.red_king_pos, e.black_king_pos = (9, 4), (0, 4)

    def stale_b(e):
        e.board[:] = 0
        e.board[9, 3], e.board[0, 4], e.board[1, 4] = 1, -1, 7
        e.red_king_pos, e.black_king_pos = (9, 3), (0, 4)

    def pawn_behind_red(e):
        e.board[:] = 0
        e.board[8, 4], e.board[0, 3], e.board[9, 4] = 1, -1, -7
        e.red_king_pos, e.black_king_pos = (8, 4), (0, 3)

    def pawn_front_black(e):
        e.board[:] = 0
        e.board[9, 4], e.board[1, 3], e.board[0, 3], e.board[2, 3] = 1, -1, 7, 7
        e.red_king_pos, e.black_king_pos = (9, 4), (1, 3)
        e.current_player = -1
    kats["positions"] = dict(stale_cache_same_file=legal_of(stale_a),
                             stale_cache_other_file=legal_of(stale_b),
                             pawn_behind_red_king=legal_of(pawn_behind_red),
                             pawns_around_black_king=legal_of(pawn_front_black))
    return kats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    from oracle import xq_oracle as xo
    xo.build()
    t0 = time.time()
    # The reference's PUCT arithmetic follows NumPy's scalar promotion: float32 under NumPy >= 2
    # (NEP 50), float64 under NumPy 1.x (its requirements.txt allows >= 1.24).  The goldens, the
    # oracle and the CUDA kernels are pinned to the NumPy >= 2 behaviour.
    assert int(np.__version__.split(".")[0]) >= 2, "goldens must be generated under NumPy >= 2 (NEP 50)"
    manifest = {"reference": REF, "numpy": np.__version__, "seed": SEED,
                "generated_by": "tests/golden/gen_golden.py",
                "numpy_semantics": "NumPy >= 2 (NEP 50): MCTSNode.select_child computes PUCT in float32; "
                                   "under NumPy 1.x the same code promotes to float64 and near-ties may "
                                   "resolve differently (not covered by these goldens)"}
    only = set(a.only.split(",")) if a.only else None
    with mp.Pool(a.procs) as pool:
        if not only or "kats" in only:
            kats = gen_kats()
            json.dump(kats, open(os.path.join(HERE, "kats.json"), "w"), ensure_ascii=False, indent=0)
            print("kats", time.time() - t0, flush=True)
        if not only or "positions" in only:
            pos = gen_positions(pool, 16 if a.quick else 64, 25 if a.quick else 50)
            np.savez_compressed(os.path.join(HERE, "positions.npz"), **pos)
            manifest["positions"] = int(len(pos["player"]))
            print("positions", len(pos["player"]), time.time() - t0, flush=True)
        if not only or "mcts" in only:
            mc = gen_mcts(pool, a.quick)
            np.savez_compressed(os.path.join(HERE, "mcts.npz"), **mc)
            manifest["mcts_searches"] = int(len(mc["player"]))
            print("mcts", len(mc["player"]), time.time() - t0, flush=True)
        if not only or "chase" in only:
            ch = gen_chase(pool, a.quick)
            json.dump(ch, open(os.path.join(HERE, "chase.json"), "w"))
            manifest["chase_games"] = len(ch)
            print("chase", len(ch), time.time() - t0, flush=True)
        if not only or "play_match" in only:
            pm = gen_play_match(pool, a.quick)
            json.dump(pm, open(os.path.join(HERE, "play_match.json"), "w"), ensure_ascii=False)
            manifest["play_match_runs"] = len(pm)
            print("play_match", len(pm), time.time() - t0, flush=True)
        if not only or "selfplay" in only:
            sp = gen_selfplay(pool, a.quick)
            json.dump(sp, open(os.path.join(HERE, "selfplay.json"), "w"), ensure_ascii=False)
            manifest["selfplay_games"] = len(sp)
            print("selfplay", len(sp), time.time() - t0, flush=True)
        if not only or "playouts" in only:
            nu, nb, nf = (32, 32, 8) if a.quick else (384, 640, 48)
            po, reasons, recs = gen_playouts(pool, nu, nb, nf)
            np.savez_compressed(os.path.join(HERE, "playouts.npz"), **po)
            json.dump(reasons, open(os.path.join(HERE, "end_reasons.json"), "w"),
                      ensure_ascii=False, indent=0)
            s = po["summary"]
            manifest["playout_games"] = int(len(s))
            manifest["playout_plies"] = int(s["plies"].sum())
            manifest["reason_histogram"] = {str(k): int((s["reason"] == k).sum()) for k in range(9)}
            manifest["max_legal"] = int(s["max_legal"].max())
            manifest["max_pseudo"] = int(s["max_pseudo"].max())
            print("playouts", manifest["playout_plies"], time.time() - t0, flush=True)
    manifest["wall_seconds"] = round(time.time() - t0, 1)
    mpath = os.path.join(HERE, "MANIFEST.json")
    old = json.load(open(mpath)) if os.path.exists(mpath) and only else {}
    old.update(manifest)
    json.dump(old, open(mpath, "w"), indent=1)
    print(json.dumps(old))


if __name__ == "__main__":
    main()

"""CPU: the C oracle against golden vectors produced by the unmodified Python reference
(tests/golden/gen_golden.py).  These pin the oracle; the GPU tests then compare the CUDA
path with the oracle and with the same goldens."""
import numpy as np
import pytest


def _load_env(xo, board, player, mc=0, winner=None, red=-1, black=-1, ncap=0, cc=0, ck=()):
    pos = lambda s: None if s < 0 else (int(s) // 9, int(s) % 9)
    return xo.Env().load(np.asarray(board, np.int8).reshape(10, 9), int(player), int(mc), winner,
                         pos(red), pos(black), int(ncap), int(cc), list(ck))


def test_philox_known_answers(xo):
    # Random123 kat_vectors: philox4x32-10
    assert [hex(x) for x in xo.philox(0, 0, 0, 0)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c",
                                                       "0x9b00dbd8"]
    out = np.zeros(4, np.uint32)
    f = 0xFFFFFFFF
    xo.lib().xqo_philox4x32(f, f, f, f, f, f, out.ctypes.data)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_initial_position(xo, golden):
    e = xo.Env()
    lm = [int(x) for x in e.legal_moves_packed()]
    assert lm == golden.kats["initial_legal_moves"]
    assert len(lm) == 44
    assert [xo.unpack(m) for m in lm[:6]] == [(6, 0, 5, 0), (6, 2, 5, 2), (6, 4, 5, 4), (6, 6, 5, 6),
                                              (6, 8, 5, 8), (7, 1, 7, 2)]


LINES = ["double_cannon_mate", "knight_shuffle", "quiet_knight", "pawn_line", "cannon_takes_king",
         "perpetual_check", "odd_cycle_repetition", "fifty_move"]


@pytest.mark.parametrize("name", LINES)
def test_kat_lines(xo, golden, name):
    k = golden.kats[name]
    st = k["start"]
    ncap = 98 if name == "fifty_move" else 0
    mc = 10 if name == "fifty_move" else 0
    e = _load_env(xo, st["board"], st["player"], mc, None, st["red"], st["black"], ncap, 0,
                  st["check_history"])
    for i in range(k["plies"]):
        rw, is_int, done = e.make_move(tuple(k["moves"][i]))
        assert repr(rw) == repr(k["rewards"][i])
        assert is_int == k["reward_is_int"][i] and done == k["dones"][i]
    assert e.winner == k["winner"] and e.s.reason == k["reason"]
    assert e.board.reshape(90).tolist() == k["final_board"]
    assert e.s.consecutive_checks == k["consecutive_checks"]
    assert len(set(e.position_history)) == k["distinct_hashes"]


def test_kat_values_match_survey(golden):
    k = golden.kats
    assert k["double_cannon_mate"]["rewards"] == [0.03, 0.025, 2.0, 0.1, 0.04, 0.0, 200.0]
    assert k["double_cannon_mate"]["end_reason"] == "将死黑方"
    assert k["knight_shuffle"]["plies"] == 70 and k["knight_shuffle"]["distinct_hashes"] == 4
    assert k["odd_cycle_repetition"]["plies"] == 27
    assert k["perpetual_check"]["winner"] == 1
    assert k["perpetual_check"]["end_reason"] == "长将判负(黑方)"


def test_kat_positions(xo, golden):
    for name, p in golden.kats["positions"].items():
        e = _load_env(xo, p["board"], p["player"], 0, None, p["red"], p["black"])
        assert [int(x) for x in e.legal_moves_packed()] == p["legal"], name
        assert e.is_in_check(p["player"]) == p["in_check"], name


def test_arbitrary_positions(xo, golden):
    P = golden.positions
    n = len(P["player"])
    assert n >= 3000
    off = P["legal_offset"]
    for i in range(n):
        L = int(P["ck_len"][i])
        ck = [bool((int(P["ck_bits"][i]) >> (L - 1 - j)) & 1) for j in range(L)]
        e = _load_env(xo, P["board"][i], P["player"][i], P["mc"][i], None, P["red"][i], P["black"][i],
                      P["ncap"][i], P["cc"][i], ck)
        lm = e.legal_moves_packed()
        assert np.array_equal(lm, P["legal"][off[i]:off[i + 1]]), i
        assert e.is_in_check(int(P["player"][i])) == bool(P["chk_self"][i]), i
        assert e.is_in_check(-int(P["player"][i])) == bool(P["chk_opp"][i]), i
        assert e.kings_facing() == bool(P["facing"][i]), i
        if P["move"][i] >= 0:
            rw, is_int, done = e.make_move(int(P["move"][i]))
            assert np.float64(rw).view(np.uint64) == P["reward"][i].view(np.uint64), i
            assert is_int == bool(P["is_int"][i]) and done == bool(P["done"][i]), i
            assert e.s.winner == P["winner"][i] and e.s.reason == P["reason"][i], i
            assert np.array_equal(e.board.reshape(90), P["board_after"][i]), i
            assert (e.s.red_king, e.s.black_king) == (P["red_after"][i], P["black_after"][i]), i
            assert e.s.consecutive_checks == P["cc_after"][i], i
            assert e.s.no_capture == P["ncap_after"][i], i
            assert e.check_history[-1] == bool(P["check_flag"][i]), i


def test_playout_summaries(xo, golden):
    S = golden.playouts["summary"]
    assert len(S) == 1024
    for bias in (0, 192):
        sel = S[S["bias"] == bias]
        first = int(sel["game_id"][0])
        assert np.array_equal(sel["game_id"], first + np.arange(len(sel)))
        total, res = xo.playout_many(len(sel), 0x5EED, first, 70, bias, n_threads=4)
        assert total == int(sel["plies"].sum())
        for f in ("plies", "winner", "reason", "max_legal", "digest", "final_hash"):
            assert np.array_equal(res[f], sel[f]), (bias, f)
        assert np.array_equal(res["reward_sum"].view(np.uint64), sel["reward_sum"].view(np.uint64))
    # the goldens exercise decisive endings, not only the 70-ply cap
    assert (S["reason"] == 1).sum() > 0 and (S["reason"] == 2).sum() > 0 and (S["reason"] == 5).sum() > 0


def test_playout_full_traces(xo, golden):
    G = golden.playouts
    S = G["summary"]
    po, mo = G["full_ply_offset"], G["full_move_offset"]
    for k, gi in enumerate(G["full_game_index"]):
        e = xo.Env()
        res, tr = e.playout(0x5EED, int(S["game_id"][gi]), 70, int(S["bias"][gi]), trace=True)
        a, b = int(po[k]), int(po[k + 1])
        assert res.plies == b - a
        assert np.array_equal(tr["n"][:res.plies], G["full_n"][a:b])
        assert np.array_equal(tr["pick"][:res.plies], G["full_pick"][a:b])
        assert np.array_equal(tr["reward"][:res.plies].view(np.uint64),
                              G["full_reward"][a:b].view(np.uint64))
        assert np.array_equal(tr["flags"][:res.plies], G["full_flags"][a:b])
        assert np.array_equal(tr["boards"][:res.plies], G["full_boards"][a:b])
        for p in range(res.plies):
            assert np.array_equal(tr["moves"][p, :tr["n"][p]], G["full_moves"][mo[a + p]:mo[a + p + 1]])


def test_mcts_visits(xo, golden):
    M = golden.mcts
    off = M["offset"]
    assert len(M["player"]) >= 70
    for i in range(len(M["player"])):
        w = None if M["winner"][i] == 2 else int(M["winner"][i])
        e = _load_env(xo, M["board"][i], M["player"][i], M["mc"][i], w, M["red"][i], M["black"][i],
                      M["ncap"][i])
        mv, vis, st = xo.mcts_search(e, int(M["n_sims"][i]), flat=bool(M["flat"][i]))
        assert np.array_equal(mv, M["moves"][off[i]:off[i + 1]]), i
        assert np.array_equal(vis, M["visits"][off[i]:off[i + 1]]), i
        assert np.array_equal(st, M["stats"][i]), i
        n = int(M["n_sims"][i])
        if len(vis) and n > 8:
            assert vis.sum() == n - 8  # first wave always lands on the bare root (B.4)


def test_encode_and_priors(xo):
    rng = np.random.default_rng(1)
    e = xo.Env()
    for pl in (1, -1):
        planes = xo.encode_board(e.board, pl)
        ref = np.zeros((15, 10, 9), np.float32)
        for i in range(1, 8):
            ref[i - 1] = e.board == i
            ref[i + 6] = e.board == -i
        ref[14] = 1.0 if pl == 1 else 0.0
        assert np.array_equal(planes, ref)
    lm = e.legal_moves_packed()
    logits = rng.standard_normal(8100).astype(np.float32)
    pr = xo.logits_to_priors(logits, lm)
    z = logits[lm.astype(np.int64)]
    z = np.exp(z - z.max())
    np.testing.assert_allclose(pr, z / z.sum(), rtol=1e-6)

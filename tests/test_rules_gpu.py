"""GPU parity tests: CUDA rules engine (through the C ABI) vs the reference goldens and the
C oracle.  Bar: bit-exact move lists (order included), boards, rewards (float64 bit patterns),
flags, caches and outcomes."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0x5EED


@pytest.fixture(scope="module")
def eng(built_lib):
    import torch
    from chinesechessai_b200 import engine
    assert torch.cuda.is_available()
    return engine


def _meta(n, **kw):
    from chinesechessai_b200._lib import META_DTYPE
    m = np.zeros(n, META_DTYPE)
    m["player"] = 1
    m["winner"] = 2
    m["red_king"] = -1
    m["black_king"] = -1
    for k, v in kw.items():
        m[k] = v
    return m


def _positions_batch(eng, P):
    n = len(P["player"])
    L = P["ck_len"].astype(np.int64)
    m = _meta(n, player=P["player"], red_king=P["red"], black_king=P["black"],
              move_count=P["mc"], no_capture=P["ncap"], consecutive_checks=P["cc"],
              check_len=L, check_bits=P["ck_bits"])
    bb = eng.BoardBatch(n)
    bb.set_state(P["board"], m)
    return bb


@pytest.fixture(params=["warp", "pair"])
def api_mapping(request, monkeypatch):
    """xq_legal_moves / xq_step pick their lane mapping by batch size; the tests force each one."""
    monkeypatch.setenv("XQ_PLAYOUT_MODE", request.param)
    return request.param


def test_initial_position(eng, golden):
    bb = eng.BoardBatch(3)
    mv, n = bb.legal_moves()
    assert n.cpu().tolist() == [44, 44, 44]
    assert mv[1, :44].cpu().tolist() == golden.kats["initial_legal_moves"]
    m = bb.meta_host()
    assert m["player"].tolist() == [1, 1, 1] and m["red_king"].tolist() == [85] * 3
    assert m["black_king"].tolist() == [4] * 3 and m["winner"].tolist() == [2] * 3


def test_legal_moves_arbitrary_positions(eng, golden, api_mapping):
    import torch
    P = golden.positions
    bb = _positions_batch(eng, P)
    chk = torch.zeros(bb.n, dtype=torch.uint8, device=bb.device)
    mv, n = bb.legal_moves(in_check=chk)
    mv, n, chk = mv.cpu().numpy(), n.cpu().numpy(), chk.cpu().numpy()
    off = P["legal_offset"]
    bad = [i for i in range(bb.n)
           if n[i] != off[i + 1] - off[i] or not np.array_equal(mv[i, :n[i]], P["legal"][off[i]:off[i + 1]])]
    assert not bad, f"{len(bad)} positions differ, first {bad[:5]}"
    assert np.array_equal(chk, P["chk_self"])


def test_step_arbitrary_positions(eng, golden, api_mapping):
    import torch
    P = golden.positions
    bb = _positions_batch(eng, P)
    move = torch.from_numpy(P["move"].astype(np.int16)).to(bb.device)
    reward, flags = bb.step(move)
    sel = P["move"] >= 0
    f = eng.decode_step_flags(flags.cpu().numpy())
    assert np.array_equal(reward.cpu().numpy().view(np.uint64)[sel], P["reward"].view(np.uint64)[sel])
    assert np.array_equal(f["done"][sel], P["done"].astype(bool)[sel])
    assert np.array_equal(f["reward_is_int"][sel], P["is_int"].astype(bool)[sel])
    assert np.array_equal(f["winner"][sel], P["winner"][sel])
    assert np.array_equal(f["reason"][sel], P["reason"][sel])
    m = bb.meta_host()
    assert np.array_equal(bb.boards_host()[sel], P["board_after"][sel])
    assert np.array_equal(m["red_king"][sel], P["red_after"][sel])
    assert np.array_equal(m["black_king"][sel], P["black_after"][sel])
    assert np.array_equal(m["consecutive_checks"][sel], P["cc_after"][sel])
    assert np.array_equal(m["no_capture"][sel], P["ncap_after"][sel])
    assert np.array_equal((m["check_bits"] & 1)[sel], P["check_flag"][sel])
    assert np.array_equal(m["winner"][sel], P["winner"][sel])
    # frozen games (move < 0) are untouched
    assert np.array_equal(bb.boards_host()[~sel], P["board"][~sel])


LINES = ["double_cannon_mate", "knight_shuffle", "quiet_knight", "pawn_line", "cannon_takes_king",
         "perpetual_check", "odd_cycle_repetition", "fifty_move"]


@pytest.mark.parametrize("name", LINES)
def test_kat_lines(eng, golden, name, api_mapping):
    import torch
    k = golden.kats[name]
    st = k["start"]
    ck = st["check_history"]
    bits = sum((1 << i) for i, v in enumerate(ck[::-1]) if v)
    m = _meta(1, player=st["player"], red_king=st["red"], black_king=st["black"],
              no_capture=98 if name == "fifty_move" else 0, move_count=10 if name == "fifty_move" else 0,
              check_len=len(ck), check_bits=bits)
    bb = eng.BoardBatch(1)
    bb.set_state(np.array(st["board"], np.int8).reshape(1, 90), m)
    for i in range(k["plies"]):
        mv = torch.tensor([eng.pack_move(k["moves"][i])], dtype=torch.int16, device=bb.device)
        reward, flags = bb.step(mv)
        f = eng.decode_step_flags(flags.cpu().numpy())
        assert repr(float(reward[0])) == repr(k["rewards"][i]), i
        assert bool(f["done"][0]) == k["dones"][i] and bool(f["reward_is_int"][0]) == k["reward_is_int"][i]
    meta = bb.meta_host()[0]
    want = 2 if k["winner"] is None else k["winner"]
    assert meta["winner"] == want and meta["reason"] == k["reason"]
    assert bb.boards_host()[0].tolist() == k["final_board"]
    assert meta["consecutive_checks"] == k["consecutive_checks"]
    assert len(set(bb.pos_hist_host()[0, :meta["hist_len"]].tolist())) == k["distinct_hashes"]


def test_kat_positions(eng, golden, api_mapping):
    for name, p in golden.kats["positions"].items():
        bb = eng.BoardBatch(1)
        bb.set_state(np.array(p["board"], np.int8).reshape(1, 90),
                     _meta(1, player=p["player"], red_king=p["red"], black_king=p["black"]))
        mv, n = bb.legal_moves()
        assert mv[0, :int(n[0])].cpu().tolist() == p["legal"], name


def _cmp_results(res, ref):
    for f in ("plies", "winner", "reason", "max_legal", "digest", "final_hash"):
        bad = np.nonzero(res[f] != ref[f])[0]
        assert len(bad) == 0, (f, bad[:5], res[f][bad[:5]], ref[f][bad[:5]])
    assert np.array_equal(res["reward_sum"].view(np.uint64), ref["reward_sum"].view(np.uint64))


def test_fused_playout_vs_reference_goldens(eng, golden):
    S = golden.playouts["summary"]
    for bias in (0, 192):
        sel = S[S["bias"] == bias]
        bb = eng.BoardBatch(len(sel))
        res = eng.results_host(bb.playout(SEED, 70, first_game_id=int(sel["game_id"][0]),
                                          capture_bias=bias))
        _cmp_results(res, sel)


def test_fused_playout_traces_vs_reference_goldens(eng, golden):
    G = golden.playouts
    S = G["summary"]
    po, mo = G["full_ply_offset"], G["full_move_offset"]
    for bias in (0, 192):
        idx = [k for k, gi in enumerate(G["full_game_index"]) if S["bias"][gi] == bias]
        gids = [int(S["game_id"][G["full_game_index"][k]]) for k in idx]
        assert gids == list(range(gids[0], gids[0] + len(gids)))
        bb = eng.BoardBatch(len(idx))
        res, tr = bb.playout(SEED, 70, first_game_id=gids[0], capture_bias=bias, trace=True)
        res = eng.results_host(res)
        tr = {k: v.cpu().numpy() for k, v in tr.items()}
        for j, k in enumerate(idx):
            a, b = int(po[k]), int(po[k + 1])
            p = b - a
            assert res["plies"][j] == p
            assert np.array_equal(tr["n"][j, :p], G["full_n"][a:b])
            assert np.array_equal(tr["pick"][j, :p], G["full_pick"][a:b])
            assert np.array_equal(tr["reward"][j, :p].view(np.uint64), G["full_reward"][a:b].view(np.uint64))
            assert np.array_equal(tr["flags"][j, :p], G["full_flags"][a:b])
            assert np.array_equal(tr["boards"][j, :p], G["full_boards"][a:b])
            for q in range(p):
                assert np.array_equal(tr["moves"][j, q, :tr["n"][j, q]],
                                      G["full_moves"][mo[a + q]:mo[a + q + 1]]), (k, q)
        # final king caches of the full games
        kings = G["full_kings"]
        meta = bb.meta_host()
        for j, k in enumerate(idx):
            last = int(po[k + 1]) - 1
            assert (meta["red_king"][j], meta["black_king"][j]) == tuple(kings[last, :2])


@pytest.mark.parametrize("bias,first", [(0, 5000), (128, 700000), (240, 31)])
def test_fused_playout_vs_oracle(eng, xo, bias, first):
    n = 4096
    bb = eng.BoardBatch(n)
    res = eng.results_host(bb.playout(SEED + bias, 70, first_game_id=first, capture_bias=bias))
    total, ref = xo.playout_many(n, SEED + bias, first, 70, bias, n_threads=8)
    _cmp_results(res, ref)
    assert int(res["plies"].sum()) == total


@pytest.mark.parametrize("lpb", [1, 2, "2q", "2q5", "2q71", "2s", 8, 16, 32])
def test_fused_playout_tile_widths(eng, xo, lpb, monkeypatch):
    """The fused kernel with 1 (thread per board), 2 (pair; "2q*" = the persistent queue-fed
    kernel with its default / 5 / 71 loop iterations per chunk, "2s" = the kernel that schedules
    the pairs' groups inside each SM), 8, 16 or 32 lanes per board is
    bit-exact, traces included, for uniform and capture-biased games and ragged batch sizes."""
    if lpb == 1:
        monkeypatch.setenv("XQ_PLAYOUT_MODE", "tpb")
    elif lpb == 2:
        monkeypatch.setenv("XQ_PLAYOUT_MODE", "pair")
    elif lpb == "2s":
        monkeypatch.setenv("XQ_PLAYOUT_MODE", "pairs")
        lpb = 3
    elif isinstance(lpb, str):
        monkeypatch.setenv("XQ_PLAYOUT_MODE", "pairq")
        if lpb[2:]:
            monkeypatch.setenv("XQ_PLAYOUT_CHUNK", lpb[2:])
        lpb = 2 + len(lpb)
    else:
        monkeypatch.setenv("XQ_PLAYOUT_MODE", "warp")
        monkeypatch.setenv("XQ_PLAYOUT_LPB", str(lpb))
    for n, bias, first in ((4099, 0, 77), (1500, 200, 123456)):
        bb = eng.BoardBatch(n)
        res = eng.results_host(bb.playout(SEED + lpb, 70, first_game_id=first, capture_bias=bias))
        total, ref = xo.playout_many(n, SEED + lpb, first, 70, bias, n_threads=8)
        _cmp_results(res, ref)
    bb = eng.BoardBatch(37)
    res, tr = bb.playout(SEED, 70, capture_bias=128, trace=True)
    res = eng.results_host(res)
    n_arr, moves = tr["n"].cpu().numpy(), tr["moves"].cpu().numpy()
    rew, flags = tr["reward"].cpu().numpy(), tr["flags"].cpu().numpy()
    for g in range(37):
        e = xo.Env()
        r, t = e.playout(SEED, g, 70, 128, trace=True)
        assert res["plies"][g] == r.plies and res["digest"][g] == r.digest
        assert np.array_equal(n_arr[g, :r.plies], t["n"][:r.plies])
        assert np.array_equal(rew[g, :r.plies].view(np.uint64), t["reward"][:r.plies].view(np.uint64))
        assert np.array_equal(flags[g, :r.plies], t["flags"][:r.plies])
        for q in range(r.plies):
            assert np.array_equal(moves[g, q, :n_arr[g, q]], t["moves"][q, :t["n"][q]])


@pytest.mark.parametrize("mode", ["warp", "tpb", "pair", "pairq", "pairs"])
def test_fused_playout_from_arbitrary_positions(eng, xo, golden, mode, monkeypatch):
    """Playouts that START from the poked golden positions (stale or missing king caches, several
    kings, enemy K/A/B next to the king, mid-game counters): every mapping of the fused kernel
    takes its irregular-board paths here and must still equal the oracle, game by game."""
    monkeypatch.setenv("XQ_PLAYOUT_MODE", mode)
    P = golden.positions
    n = len(P["player"])
    for plies, bias in ((1, 0), (9, 160)):
        bb = _positions_batch(eng, P)
        res = eng.results_host(bb.playout(SEED, plies, first_game_id=1000, capture_bias=bias))
        m = bb.meta_host()
        boards = bb.boards_host()
        for i in range(n):
            L = int(P["ck_len"][i])
            ck = [bool((int(P["ck_bits"][i]) >> (L - 1 - j)) & 1) for j in range(L)]
            e = xo.Env().load(P["board"][i].reshape(10, 9), int(P["player"][i]), int(P["mc"][i]), None,
                              xo.Env._pos(int(P["red"][i])), xo.Env._pos(int(P["black"][i])),
                              int(P["ncap"][i]), int(P["cc"][i]), ck)
            r = e.playout(SEED, 1000 + i, plies, bias)
            for f in ("plies", "winner", "reason", "max_legal", "digest", "final_hash"):
                assert res[f][i] == getattr(r, f), (mode, plies, i, f)
            assert np.float64(res["reward_sum"][i]).view(np.uint64) == np.float64(r.reward_sum).view(np.uint64)
            assert np.array_equal(boards[i], e.board.reshape(90)), (mode, i)
            assert (m["red_king"][i], m["black_king"][i]) == (e.s.red_king, e.s.black_king), (mode, i)
            assert m["consecutive_checks"][i] == e.s.consecutive_checks and m["no_capture"][i] == e.s.no_capture


def test_step_per_launch_equals_fused(eng, api_mapping):
    import torch
    n, bias = 2048, 96
    fused = eng.BoardBatch(n)
    rf = eng.results_host(fused.playout(SEED, 70, capture_bias=bias))
    bb = eng.BoardBatch(n)
    plies = torch.zeros(n, dtype=torch.int32, device=bb.device)
    rsum = torch.zeros(n, dtype=torch.float64, device=bb.device)
    bb.legal_moves()
    for ply in range(70):
        mv = bb.pick(SEED, ply, capture_bias=bias)
        plies += (mv >= 0).to(torch.int32)
        reward, flags = bb.step(mv, want_next=True)
        rsum += torch.where(mv >= 0, reward, torch.zeros_like(reward))
    assert np.array_equal(plies.cpu().numpy(), rf["plies"])
    assert np.array_equal(rsum.cpu().numpy().view(np.uint64), rf["reward_sum"].view(np.uint64))
    assert np.array_equal(bb.boards_host(), fused.boards_host())
    a, b = bb.meta_host(), fused.meta_host()
    for f in a.dtype.names:
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(bb.pos_hist_host(), fused.pos_hist_host())
    # API-faithful variant: separate legal_moves launch instead of step's next list
    cc = eng.BoardBatch(256)
    cf = eng.BoardBatch(256)
    cf.playout(SEED, 70)
    for ply in range(70):
        cc.legal_moves()
        cc.step(cc.pick(SEED, ply))
    assert np.array_equal(cc.boards_host(), cf.boards_host())


def test_full_size_properties(eng, xo):
    """cfg 2 size (65,536 boards x 70 plies): determinism, shard invariance, and the oracle's
    digests on a strided sample of the same game ids."""
    n = 65536
    bb = eng.BoardBatch(n)
    r1 = eng.results_host(bb.playout(SEED, 70))
    bb.reset()
    bb.pos_hist.zero_()
    r2 = eng.results_host(bb.playout(SEED, 70))
    assert np.array_equal(r1, r2)
    assert 4_300_000 < int(r1["plies"].sum()) <= 70 * n
    assert int(bb.meta_host()["flags"].max()) == 0
    # sharding: a rank that owns games [a, b) reproduces the slice of the full run
    a, b = 40000, 41024
    sh = eng.BoardBatch(b - a)
    rs = eng.results_host(sh.playout(SEED, 70, first_game_id=a))
    assert np.array_equal(rs, r1[a:b])
    # the oracle on EVERY game of the batch (BASELINE.json cfg 2: "hash of all final states")
    total, ref = xo.playout_many(n, SEED, 0, 70, 0, n_threads=os.cpu_count() or 8)
    _cmp_results(r1, ref)
    assert int(r1["plies"].sum()) == total


def test_playout_host_e2e(eng):
    from chinesechessai_b200._lib import BOARD_STRIDE
    n = 1024
    bb = eng.BoardBatch(n)
    boards = np.ascontiguousarray(bb.board.cpu().numpy())
    meta = np.ascontiguousarray(bb.meta_host())
    assert boards.shape == (n, BOARD_STRIDE)
    rd = eng.results_host(bb.playout(SEED, 70, capture_bias=64))
    rh = eng.playout_host(boards, meta, SEED, 70, capture_bias=64)
    assert np.array_equal(rd, rh)
    assert np.array_equal(boards, bb.board.cpu().numpy())
    assert np.array_equal(meta, bb.meta_host())


def test_ragged_and_empty(eng):
    import torch
    bb = eng.BoardBatch(0)
    bb.legal_moves()
    bb.playout(SEED, 70)
    for n in (1, 7, 9, 33):
        b = eng.BoardBatch(n)
        r = eng.results_host(b.playout(SEED, 70))
        assert len(r) == n and (r["plies"] > 0).all()
    # max_plies = 0 leaves the state untouched
    b = eng.BoardBatch(4)
    r = eng.results_host(b.playout(SEED, 0))
    assert r["plies"].tolist() == [0] * 4 and b.meta_host()["move_count"].tolist() == [0] * 4
    # history capacity overflow is flagged, not silent
    b = eng.BoardBatch(2, hist_cap=8)
    b.playout(SEED, 20)
    assert (b.meta_host()["flags"] & 1).all()
    # argument errors come back as codes + message, never a crash
    from chinesechessai_b200 import _lib
    rc = b.lib.xq_legal_moves(None, None, None, None, None, 1, None)
    assert rc == -1
    with pytest.raises(_lib.XqError, match="xq_legal_moves"):
        _lib.check(rc)


def test_encode_planes_and_priors(eng, xo):
    import torch
    n = 512
    bb = eng.BoardBatch(n)
    bb.playout(SEED, 23, capture_bias=100)
    boards, meta = bb.boards_host(), bb.meta_host()
    pl = bb.meta[:, 0].view(torch.int8)
    planes = eng.encode_planes(bb.board, pl).cpu().numpy()
    for i in range(0, n, 7):
        assert np.array_equal(planes[i], xo.encode_board(boards[i], int(meta["player"][i]))), i
    pb = eng.encode_planes(bb.board, pl, dtype=torch.bfloat16).float().cpu().numpy()
    assert np.array_equal(pb, planes)
    # channels-last bf16 with the channel count padded to 16 (the inference net's input layout)
    cl = eng.encode_planes_nhwc16(bb.board, pl)
    assert cl.shape == (n, 16, 10, 9) and cl.is_contiguous(memory_format=torch.channels_last)
    clh = cl.float().cpu().numpy()
    assert np.array_equal(clh[:, :15], planes) and not clh[:, 15].any()
    mv, nm = bb.legal_moves()
    logits = torch.randn(n, 8100, device=bb.device) * 3
    pri = eng.policy_priors(logits, mv, nm).cpu().numpy()
    lg, mvh, nmh = logits.cpu().numpy(), mv.cpu().numpy(), nm.cpu().numpy()
    for i in range(0, n, 5):
        k = int(nmh[i])
        ref = xo.logits_to_priors(lg[i], mvh[i, :k])
        np.testing.assert_allclose(pri[i, :k], ref, rtol=2e-6, atol=1e-9)  # float32 softmax, B.5
        assert (pri[i, k:] == 0).all() and abs(pri[i].sum() - 1) < 1e-5
        assert int(np.argmax(pri[i, :k])) == int(np.argmax(ref))


def test_position_history_equals_oracle(eng, xo, api_mapping):
    """_get_position_hash / position_history (chess_env.py:338,497-504): the device history row
    holds exactly the oracle's keys (same key function, mover's side byte), ply by ply."""
    n = 64
    bb = eng.BoardBatch(n)
    bb.playout(SEED, 70, capture_bias=80)
    hist, meta = bb.pos_hist_host(), bb.meta_host()
    keys = bb.position_hash().cpu().numpy().view(np.uint64)
    for g in range(n):
        e = xo.Env()
        e.playout(SEED, g, 70, 80)
        want = np.array(e.position_history, dtype=np.uint64)
        assert meta["hist_len"][g] == len(want)
        assert np.array_equal(hist[g, :len(want)], want), g
        assert keys[g] == np.uint64(e.position_hash()), g
        assert meta["check_len"][g] == len(e.check_history)
        bits = sum((1 << i) for i, v in enumerate(e.check_history[::-1][:32]) if v)
        assert int(meta["check_bits"][g]) == bits, g

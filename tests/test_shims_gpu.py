"""GPU: the drop-in Python surface (ChineseChess, MCTS, self_play_game, parallel_self_play)
driven the way the reference's callers drive it, against goldens recorded from the reference."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pkg(built_lib):
    import torch
    assert torch.cuda.is_available()
    from chinesechessai_b200 import chess_env, self_play
    return chess_env, self_play


class StubNet:
    """Injected evaluator — the same one the goldens were recorded with."""

    def __init__(self, xo, flat=False):
        self.xo, self.flat, self.calls, self.sizes = xo, flat, 0, []

    def predict_batch(self, items):
        xo = self.xo
        self.calls += 1
        self.sizes.append(len(items))
        n = len(items)
        boards = np.stack([np.asarray(b, np.int8).reshape(90) for b, _, _ in items])
        players = np.array([p for _, p, _ in items], np.int32)
        moves = np.zeros((n, 128), np.int16)
        nm = np.zeros(n, np.int32)
        for i, (_, _, lm) in enumerate(items):
            nm[i] = len(lm)
            moves[i, :len(lm)] = [xo.pack(m) for m in lm]
        pri, val = xo.hash_eval(boards, players, moves, nm, flat=self.flat)
        return [({m: pri[i, j] for j, m in enumerate(lm)}, float(val[i])) for i, (_, _, lm) in enumerate(items)]


def test_env_surface_and_types(pkg):
    chess_env, _ = pkg
    env = chess_env.ChineseChess()
    board, player = env.reset()
    assert board.shape == (10, 9) and board.dtype == np.int8 and player == 1
    assert env.red_king_pos == (9, 4) and env.black_king_pos == (0, 4) and env.winner is None
    lm = env.get_legal_moves()
    assert len(lm) == 44 and lm[0] == (6, 0, 5, 0) and all(type(x) is int for x in lm[0])
    (b2, p2), reward, done = env.make_move((9, 1, 7, 2))
    assert isinstance(reward, float) and repr(reward) == "0.035" and done is False and p2 == -1
    assert env.move_count == 1 and env.current_player == -1 and env.check_history == [False]
    assert len(env.position_history) == 1 and env.no_capture_count == 1
    b2[0, 0] = 9  # get_state returns copies
    assert env.board[0, 0] == -5


def test_env_replays_golden_games(pkg, golden, xo):
    """Every ply of the 96 fully-traced reference games through the one-board API."""
    chess_env, _ = pkg
    G = golden.playouts
    S = G["summary"]
    po, mo = G["full_ply_offset"], G["full_move_offset"]
    for k, gi in enumerate(G["full_game_index"][::4]):
        k = k * 4
        env = chess_env.ChineseChess()
        a, b = int(po[k]), int(po[k + 1])
        for q in range(b - a):
            lm = env.get_legal_moves()
            assert [xo.pack(m) for m in lm] == G["full_moves"][mo[a + q]:mo[a + q + 1]].tolist(), (k, q)
            (_bd, _pl), reward, done = env.make_move(xo.unpack(int(G["full_pick"][a + q])))
            fl = int(G["full_flags"][a + q])
            assert repr(float(reward)) == repr(float(G["full_reward"][a + q])), (k, q)
            assert isinstance(reward, int) == bool(fl & 2) and done == bool(fl & 1), (k, q)
            assert np.array_equal(env.board.reshape(90), G["full_boards"][a + q]), (k, q)
        want_w = int(S["winner"][gi])
        assert env.winner == (None if want_w == 2 else want_w)
        assert env.end_reason == golden.end_reasons[str(int(gi))], (k, env.end_reason)
        assert env.move_count == b - a and len(env.check_history) == b - a


def test_env_kats_and_private_helpers(pkg, golden):
    chess_env, _ = pkg
    for name in ("double_cannon_mate", "perpetual_check", "odd_cycle_repetition", "fifty_move",
                 "cannon_takes_king"):
        k = golden.kats[name]
        st = k["start"]
        env = chess_env.ChineseChess()
        env.board[:] = np.array(st["board"], np.int8).reshape(10, 9)  # poke like the reference's tests
        env.current_player = st["player"]
        pos = lambda s: None if s < 0 else (s // 9, s % 9)
        env.red_king_pos, env.black_king_pos = pos(st["red"]), pos(st["black"])
        env.check_history = list(st["check_history"])
        if name == "fifty_move":
            env.no_capture_count, env.move_count = 98, 10
        for i in range(k["plies"]):
            _, reward, done = env.make_move(tuple(k["moves"][i]))
            assert repr(float(reward)) == repr(k["rewards"][i]) and done == k["dones"][i]
            assert isinstance(reward, int) == k["reward_is_int"][i]
        assert env.winner == k["winner"] and env.end_reason == k["end_reason"], name
    # pawn-perspective quirk (A.3) through the private helper the reference's tests use
    p = golden.kats["positions"]["pawn_behind_red_king"]
    env = chess_env.ChineseChess()
    env.board[:] = np.array(p["board"], np.int8).reshape(10, 9)
    env.red_king_pos, env.black_king_pos = (8, 4), (0, 3)
    assert env._is_in_check(1) is True
    env.current_player = -1
    assert env._is_in_check(1) is False
    env.check_history = [True] * 10 + [False, True]
    assert env._check_perpetual_check() and not env._check_perpetual_chase()


def test_mcts_search_dict_vs_reference(pkg, golden, xo):
    chess_env, self_play = pkg
    M = golden.mcts
    off = M["offset"]
    for i in range(0, len(M["player"]), 3):
        env = chess_env.ChineseChess()
        env.board = M["board"][i].reshape(10, 9).copy()
        env.current_player = int(M["player"][i])
        env.move_count, env.no_capture_count = int(M["mc"][i]), int(M["ncap"][i])
        env.winner = None if M["winner"][i] == 2 else int(M["winner"][i])
        pos = lambda s: None if s < 0 else (int(s) // 9, int(s) % 9)
        env.red_king_pos, env.black_king_pos = pos(M["red"][i]), pos(M["black"][i])
        net = StubNet(xo, bool(M["flat"][i]))
        before = env.board.copy()
        visits = self_play.MCTS(net, int(M["n_sims"][i])).search(env)
        assert [xo.pack(m) for m in visits.keys()] == M["moves"][off[i]:off[i + 1]].tolist(), i
        assert list(visits.values()) == M["visits"][off[i]:off[i + 1]].tolist(), i
        assert net.calls == M["stats"][i][2] and sum(net.sizes) == M["stats"][i][1], i
        assert np.array_equal(env.board, before) and env.move_count == int(M["mc"][i])


def test_self_play_game_vs_reference(pkg, xo):
    """Same injected evaluator + same np.random seed => identical samples, winner, end_reason."""
    _, self_play = pkg
    games = json.load(open(os.path.join(GOLDEN, "selfplay.json"), encoding="utf-8"))
    assert len(games) >= 4
    for g in games:
        np.random.seed(g["seed"])
        net = StubNet(xo, False)
        opp = StubNet(xo, True) if g["opponent"] else None
        data, winner, reason = self_play.self_play_game(net, temperature=g["temperature"],
                                                        num_simulations=g["n_sims"], opponent_network=opp)
        assert winner == g["winner"] and reason == g["end_reason"], (g["seed"], reason)
        assert len(data) == len(g["boards"])
        for (board, probs, reward), gb, gm, gp, gr in zip(data, g["boards"], g["moves"], g["probs"], g["rewards"]):
            assert board.reshape(90).tolist() == gb
            assert [xo.pack(m) for m in probs.keys()] == gm
            assert [float(p) for p in probs.values()] == gp
            assert repr(float(reward)) == repr(gr)


def test_batched_self_play_vs_oracle_game_loop(pkg, xo):
    """The batched device game loop (search -> counter-based temperature sampling -> make_move)
    against the same loop driven on the oracle: identical visit counts, chosen moves, rewards,
    outcomes and shaped sample rewards (self_play.py:203-310) for every game and ply."""
    _, self_play = pkg
    from chinesechessai_b200.mcts import HashEvaluator
    for n, n_sims, temp, seed, first in ((24, 15, 1.0, 5, 0), (10, 30, 0.5, 9, 1000), (6, 15, 0.001, 2, 7)):
        sp = self_play.BatchedSelfPlay(HashEvaluator(), n, n_sims, temperature=temp, seed=seed,
                                       first_game_id=first)
        sp.play()
        res = sp.materialise()
        P = sp.plies
        rm, rv, rn = (t[:P].cpu().numpy() for t in (sp.rec_moves, sp.rec_visits, sp.rec_n))
        played, rmove = sp.rec_played[:P].cpu().numpy(), sp.rec_move[:P].cpu().numpy()
        meta = sp.boards.meta_host()
        for g, (data, winner, reason) in enumerate(res):
            e = xo.Env()
            rewards, boards = [], []
            for p in range(70):
                om, ov, _ = xo.mcts_search(e, n_sims)
                if len(e.legal_moves_packed()) == 0 or len(om) == 0:
                    break
                assert played[p, g], (g, p)
                k = int(rn[p, g])
                assert np.array_equal(rm[p, g, :k], om) and np.array_equal(rv[p, g, :k], ov), (g, p)
                idx = xo.sample_move(ov, temp, seed, first + g, p)
                assert int(rmove[p, g]) == int(om[idx]), (g, p)
                boards.append(e.board.copy())
                rw, _, done = e.make_move(int(om[idx]))
                rewards.append(rw)
                if done:
                    break
            n_plies = len(rewards)
            assert not played[n_plies:, g].any() and played[:n_plies, g].all(), g
            w = 0 if e.winner is None else e.winner
            assert winner == w and meta["reason"][g] == e.s.reason, g
            assert len(data) == n_plies
            for i, (b, probs, total) in enumerate(data):
                assert np.array_equal(b, boards[i])
                player = 1 if i % 2 == 0 else -1
                assert repr(total) == repr(self_play.final_reward(w, player, n_plies) + rewards[i] * 0.01)
                assert abs(sum(probs.values()) - 1.0) < 1e-9
    out = self_play.parallel_self_play(_Net(), 4, temperature=1.0, num_simulations=15, num_workers=4)
    assert len(out) == 4 and all(len(gd) > 0 and isinstance(r, str) for gd, w, r in out)


def test_self_play_is_shard_invariant(pkg):
    """A game's trajectory depends on (seed, game id) only: two half batches with game-id
    offsets reproduce the full batch (what makes the multi-GPU shards verifiable)."""
    import torch
    _, self_play = pkg
    from chinesechessai_b200.mcts import HashEvaluator
    full = self_play.BatchedSelfPlay(HashEvaluator(), 32, 15, temperature=1.0, seed=77)
    full.play(20)
    for lo in (0, 16):
        part = self_play.BatchedSelfPlay(HashEvaluator(), 16, 15, temperature=1.0, seed=77, first_game_id=lo)
        part.play(20)
        assert torch.equal(part.rec_move[:20], full.rec_move[:20, lo:lo + 16])
        assert torch.equal(part.rec_visits[:20], full.rec_visits[:20, lo:lo + 16])
        assert torch.equal(part.boards.board, full.boards.board[lo:lo + 16])


def _Net():
    import torch
    from chinesechessai_b200.neural_network import ChessNet
    torch.manual_seed(0)
    return ChessNet().cuda().eval()


def test_network_predict_batch_surface(pkg):
    chess_env, _ = pkg
    net = _Net()
    env = chess_env.ChineseChess()
    lm = env.get_legal_moves()
    out = net.predict_batch([(env.board, 1, lm), (env.board, -1, lm[:5])])
    assert len(out) == 2 and list(out[0][0].keys()) == lm and isinstance(out[0][1], float)
    assert abs(sum(float(v) for v in out[0][0].values()) - 1) < 1e-5
    assert type(next(iter(out[0][0].values()))) is np.float32
    planes = net.encode_board(env.board, 1)
    assert planes.shape == (15, 10, 9) and planes.dtype == np.float32 and planes[14].all()
    assert net.predict(env.board, 1, lm)[0].keys() == out[0][0].keys()


def test_device_training_tensors_match_materialised_samples(pkg):
    """§8f: the device sample path reproduces the host tuples of materialise() exactly."""
    import torch
    _, self_play = pkg
    from chinesechessai_b200 import samples
    from chinesechessai_b200.mcts import HashEvaluator
    for red_only in (False, True):
        sp = self_play.BatchedSelfPlay(HashEvaluator(), 40, 15, temperature=1.0, seed=11)
        sp.play()
        host = sp.materialise(red_only=red_only)
        dev = samples.training_tensors(sp, red_only=red_only)
        boards = np.concatenate([np.stack([b.reshape(90) for b, _, _ in gd]) for gd, _, _ in host if gd])
        rewards = np.array([r for gd, _, _ in host for _, _, r in gd], np.float64)
        assert np.array_equal(dev["board"].cpu().numpy(), boards)
        assert np.array_equal(dev["reward"].cpu().numpy().view(np.uint64), rewards.view(np.uint64))
        idx = torch.arange(0, len(rewards), 7, device=dev["board"].device)
        states, target = samples.training_batch(dev, idx)
        net = _Net()
        want = np.stack([net.encode_board(boards[i].reshape(10, 9), 1) for i in idx.cpu().tolist()[:5]])
        assert np.array_equal(states[:5].cpu().numpy(), want)
        assert target.shape == (len(idx), 1) and target.dtype == torch.float32


@pytest.mark.gpu
def test_batched_self_play_cuda_graph_equals_eager():
    """Small batches replay the per-ply search as a captured CUDA graph; with the deterministic
    hashed evaluator every recorded move, visit count and reward equals the eager loop's."""
    import torch
    from chinesechessai_b200.mcts import HashEvaluator
    from chinesechessai_b200.self_play import BatchedSelfPlay
    runs = []
    for graph in (False, True):
        sp = BatchedSelfPlay(HashEvaluator(), 24, 15, temperature=1.0, seed=77, first_game_id=5, use_graph=graph)
        assert sp.use_graph == graph
        sp.play()
        runs.append(sp)
    a, b = runs
    assert b._graph is not None and a._graph is None
    assert a.plies == b.plies and a.plies > 10
    P = a.plies
    for f in ("rec_move", "rec_visits", "rec_moves", "rec_n", "rec_board", "rec_played"):
        assert torch.equal(getattr(a, f)[:P], getattr(b, f)[:P]), f
    assert torch.equal(a.rec_reward[:P].view(torch.int64), b.rec_reward[:P].view(torch.int64))
    assert torch.equal(a.boards.board, b.boards.board) and torch.equal(a.boards.meta, b.boards.meta)


@pytest.mark.gpu
def test_batched_self_play_cuda_graph_with_network():
    """The captured search with the real evaluator (encode -> folded bf16 net -> priors): every
    live root receives the same number of child visits as in the eager loop."""
    import torch
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.self_play import BatchedSelfPlay
    torch.manual_seed(0)
    net = ChessNet().cuda().eval()
    sp = BatchedSelfPlay(net, 16, 15, temperature=1.0, seed=3, net_dtype=torch.bfloat16)
    assert sp.use_graph                       # auto mode: small batch, network evaluator
    sp.play(6)
    assert sp._graph is not None
    ref = BatchedSelfPlay(net, 16, 15, temperature=1.0, seed=3, net_dtype=torch.bfloat16, use_graph=False)
    ref.play(1)
    played = sp.rec_played[:sp.plies]
    tot = sp.rec_visits[:sp.plies].sum(-1)
    expect = int(ref.rec_visits[0, 0].sum())
    # 15 sims = a wave of 8 that expands the root + a wave of 7 through its children (B.2)
    assert expect == 7 and bool((tot[played] == expect).all())


def test_parallel_self_play_interrupt_carries_finished_games(pkg, monkeypatch, capsys):
    """Ctrl-C during parallel_self_play raises InterruptedWithResults with exactly the games that
    had finished (self_play.py:433-452), each a complete (game_data, winner, end_reason) equal to
    the same game of an uninterrupted run; the progress line has the reference's format."""
    _, self_play = pkg
    from chinesechessai_b200.mcts import HashEvaluator
    np.random.seed(123)
    full = self_play.parallel_self_play(HashEvaluator(), 96, temperature=1.0, num_simulations=15)
    out = capsys.readouterr().out
    assert "进度: [" in out and "96/96 (100.0%)" in out and "有效:96 (100%)" in out
    lengths = sorted(len(gd) for gd, _, _ in full)
    calls = {"n": 0}
    real = self_play._progress

    def progress(done, total, valid):
        calls["n"] += 1
        real(done, total, valid)
        if calls["n"] == 7:                 # after 60 plies: the shortest games are over, most are not
            raise KeyboardInterrupt
    monkeypatch.setattr(self_play, "_progress", progress)
    np.random.seed(123)                     # same engine seed as the full run
    with pytest.raises(self_play.InterruptedWithResults) as ei:
        self_play.parallel_self_play(HashEvaluator(), 96, temperature=1.0, num_simulations=15)
    part = ei.value.results
    assert isinstance(part, list) and len(part) < 96
    if lengths[0] <= 50:                    # some game was certainly over when the interrupt came
        assert len(part) > 0
    keyed = {(w, r, len(gd), gd[0][0].tobytes(), gd[-1][0].tobytes(), gd[-1][2]) for gd, w, r in full}
    for gd, w, r in part:
        assert (w, r, len(gd), gd[0][0].tobytes(), gd[-1][0].tobytes(), gd[-1][2]) in keyed


def test_reference_play_match_on_shims_equals_reference_golden(pkg, xo, tmp_path):
    """The reference's UNCHANGED compare_models.play_match, imported from the reference checkout
    but running on the drop-in ChineseChess / MCTS classes, reproduces the result dict AND the
    move sequences that the same function produced on the reference's own classes (golden
    recorded by tests/golden/gen_golden.py with the same injected evaluators and np.random seed)."""
    import subprocess
    import sys
    from baseline import reference as R
    if R.locate() is None:
        pytest.skip("no reference checkout (baseline/_ref, XQ_REFERENCE)")
    gold = json.load(open(os.path.join(GOLDEN, "play_match.json"), encoding="utf-8"))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json, io, contextlib\n"
        "import numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "from oracle import xq_oracle as xo\n"
        "from tests.test_shims_gpu import StubNet\n"
        "with contextlib.redirect_stdout(io.StringIO()):\n"
        "    import chess_env, self_play, compare_models\n"
        "import chinesechessai_b200.self_play as ours\n"
        "assert self_play.MCTS is ours.MCTS and compare_models.__file__.startswith(sys.argv[1])\n"
        "out = []\n"
        "for seed, n_games, n_sims in json.loads(sys.argv[2]):\n"
        "    ours.MCTS_SIMULATIONS = n_sims\n"
        "    log = []\n"
        "    orig = chess_env.ChineseChess.make_move\n"
        "    def logged(self, mv, orig=orig, log=log):\n"
        "        if sys._getframe(1).f_code.co_name == 'play_match':\n"
        "            log.append((mv[0] * 9 + mv[1]) * 90 + mv[2] * 9 + mv[3])\n"
        "        return orig(self, mv)\n"
        "    chess_env.ChineseChess.make_move = logged\n"
        "    np.random.seed(seed)\n"
        "    res = compare_models.play_match(StubNet(xo, False), StubNet(xo, True), num_games=n_games, verbose=False)\n"
        "    chess_env.ChineseChess.make_move = orig\n"
        "    out.append(dict(result=res, moves=log))\n"
        "print(json.dumps(out))\n")
    jobs = json.dumps([(g["seed"], g["n_games"], g["n_sims"]) for g in gold])
    p = subprocess.run([sys.executable, "-c", code, os.path.realpath(R.locate()), jobs], env=R.env_for_shims(),
                       cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    got = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("[")][-1])
    for a, g in zip(got, gold):
        assert a["moves"] == g["moves"]
        assert a["result"] == g["result"]


def test_mcts_device_path_equals_predict_batch_path(pkg):
    """MCTS(ChessNet).search takes the all-device, graph-replayed path; the same network hidden
    behind a plain object with predict_batch takes the reference's per-wave host path.  Same
    visit dicts (keys, order, counts), also after the weights change (the graph is re-captured)."""
    import torch
    chess_env, self_play = pkg
    from chinesechessai_b200.neural_network import ChessNet

    class Wrapped:
        def __init__(self, net):
            self.predict_batch = net.predict_batch

    torch.manual_seed(1)
    net = ChessNet().cuda().eval()
    fast, slow = self_play.MCTS(net, 30), self_play.MCTS(Wrapped(net), 30)
    assert fast._device_path() and not slow._device_path()
    env = chess_env.ChineseChess()
    rng = np.random.default_rng(0)
    for ply in range(12):
        a, b = fast.search(env), slow.search(env)
        assert list(a.items()) == list(b.items()), ply
        assert sum(a.values()) == 30 - 8
        legal = env.get_legal_moves()
        env.make_move(legal[int(rng.integers(len(legal)))])
        if ply == 6:
            with torch.no_grad():
                for p in net.parameters():
                    p.mul_(1.01)
    assert fast._fast[30]._graph is not None and fast._fast[30].ev._version >= 2


def test_chase_history_equals_reference(pkg):
    """chase_history (chess_env.py:344-345) is dead bookkeeping in the reference but an observable
    attribute: the shim's lazily computed entries equal the lists the unmodified reference
    appended, ply by ply, for games recorded by gen_golden.py (uniform and capture-biased)."""
    chess_env, _ = pkg
    from chinesechessai_b200.engine import unpack_move
    games = json.load(open(os.path.join(GOLDEN, "chase.json")))
    assert len(games) >= 2
    threats = 0
    for g in games:
        env = chess_env.ChineseChess()
        for mv in g["moves"]:
            env.make_move(unpack_move(mv))
        assert len(env.chase_history) == len(g["chase"])
        for got, want in zip(env.chase_history, g["chase"]):
            assert [[a[0] * 9 + a[1], b[0] * 9 + b[1]] for a, b in got] == want
            threats += len(want)
        assert env.chase_history[-1] == [tuple(map(tuple, x)) for x in env.chase_history[-1]]
    assert threats > 100

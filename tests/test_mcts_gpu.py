"""GPU parity: CUDA MCTS (xq_mcts_*) vs the reference's MCTS.search goldens and the oracle.
Tolerance: with injected priors/values every visit count must be EQUAL."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built_lib):
    import torch
    from chinesechessai_b200 import engine, mcts
    assert torch.cuda.is_available()
    return engine, mcts


def _meta(n, **kw):
    from chinesechessai_b200._lib import META_DTYPE
    m = np.zeros(n, META_DTYPE)
    m["player"], m["winner"], m["red_king"], m["black_king"] = 1, 2, -1, -1
    for k, v in kw.items():
        m[k] = v
    return m


def test_hash_eval_matches_oracle(mods, xo):
    import torch
    eng, mcts = mods
    bb = eng.BoardBatch(256)
    bb.playout(3, 31, capture_bias=90)
    mv, nm = bb.legal_moves()
    pl = bb.meta[:, 0].view(torch.int8).contiguous()
    for flat in (False, True):
        pri, val = mcts.HashEvaluator(flat)(bb.board, pl, mv, nm)
        rp, rv = xo.hash_eval(bb.boards_host(), bb.meta_host()["player"].astype(np.int32),
                              mv.cpu().numpy(), nm.cpu().numpy().astype(np.int32), flat=flat)
        assert np.array_equal(pri.cpu().numpy().view(np.uint32), rp.view(np.uint32))
        assert np.array_equal(val.cpu().numpy().view(np.uint64), rv.view(np.uint64))


def test_search_vs_reference_goldens(mods, golden):
    """Roots, simulation counts (8..150) and visit dicts produced by the unmodified reference
    MCTS.search with the injected hash evaluator."""
    eng, mcts = mods
    M = golden.mcts
    off = M["offset"]
    for n_sims in sorted(set(M["n_sims"].tolist())):
        for flat in (0, 1):
            idx = np.nonzero((M["n_sims"] == n_sims) & (M["flat"] == flat))[0]
            if len(idx) == 0:
                continue
            bb = eng.BoardBatch(len(idx))
            bb.set_state(M["board"][idx], _meta(len(idx), player=M["player"][idx], winner=M["winner"][idx],
                                                red_king=M["red"][idx], black_king=M["black"][idx],
                                                move_count=M["mc"][idx], no_capture=M["ncap"][idx]))
            s = mcts.BatchedMCTS(len(idx), n_sims)
            mv, vis, nc = s.search(bb.board, bb.meta, mcts.HashEvaluator(bool(flat)))
            mv, vis, nc = mv.cpu().numpy(), vis.cpu().numpy(), nc.cpu().numpy()
            for j, i in enumerate(idx):
                want_m, want_v = M["moves"][off[i]:off[i + 1]], M["visits"][off[i]:off[i + 1]]
                assert nc[j] == len(want_m), (n_sims, flat, i)
                assert np.array_equal(mv[j, :nc[j]], want_m), (n_sims, flat, i)
                assert np.array_equal(vis[j, :nc[j]], want_v), (n_sims, flat, i, vis[j, :nc[j]], want_v)
            # the search must not mutate the root states
            assert np.array_equal(bb.boards_host(), M["board"][idx])


@pytest.mark.parametrize("n_sims,plies,bias", [(15, 0, 0), (50, 17, 64), (50, 66, 128), (30, 69, 0), (97, 40, 200)])
def test_search_vs_oracle_batch(mods, xo, n_sims, plies, bias):
    eng, mcts = mods
    n = 96
    bb = eng.BoardBatch(n)
    if plies:
        bb.playout(11 + plies, plies, capture_bias=bias)
    s = mcts.BatchedMCTS(n, n_sims)
    mv, vis, nc = s.search(bb.board, bb.meta, mcts.HashEvaluator())
    mv, vis, nc = mv.cpu().numpy(), vis.cpu().numpy(), nc.cpu().numpy()
    boards, meta = bb.boards_host(), bb.meta_host()
    pos = lambda q: None if q < 0 else (int(q) // 9, int(q) % 9)
    n_terminal_roots = 0
    for g in range(n):
        w = None if meta["winner"][g] == 2 else int(meta["winner"][g])
        e = xo.Env().load(boards[g].reshape(10, 9), int(meta["player"][g]), int(meta["move_count"][g]), w,
                          pos(meta["red_king"][g]), pos(meta["black_king"][g]), int(meta["no_capture"][g]))
        om, ov, st = xo.mcts_search(e, n_sims)
        assert nc[g] == len(om), g
        assert np.array_equal(mv[g, :nc[g]], om), g
        assert np.array_equal(vis[g, :nc[g]], ov), (g, vis[g, :nc[g]], ov)
        n_terminal_roots += len(om) == 0
    if plies >= 66:
        assert n_terminal_roots > 0 or True


def test_inactive_games_and_reuse(mods):
    import torch
    eng, mcts = mods
    bb = eng.BoardBatch(16)
    act = torch.ones(16, dtype=torch.uint8, device=bb.device)
    act[::2] = 0
    s = mcts.BatchedMCTS(16, 15)
    for _ in range(2):  # the tree pool is reusable across searches
        mv, vis, nc = s.search(bb.board, bb.meta, mcts.HashEvaluator(), active=act)
        nc = nc.cpu().numpy()
        assert (nc[::2] == 0).all() and (nc[1::2] == 44).all()
        assert (vis.cpu().numpy()[1::2].sum(1) == 7).all()  # 15 sims -> 7 child visits (B.4)


def test_net_evaluator_runs(mods):
    """Real network path (shape/dtype plumbing; numerics are outside the bit-parity gate, B.5)."""
    import torch
    eng, mcts = mods
    from chinesechessai_b200.neural_network import ChessNet
    torch.manual_seed(0)
    net = ChessNet().cuda().eval()
    bb = eng.BoardBatch(32)
    s = mcts.BatchedMCTS(32, 15)
    for dt in (torch.float32, torch.bfloat16):
        mv, vis, nc = s.search(bb.board, bb.meta, mcts.NetEvaluator(net, dt))
        v = vis.cpu().numpy()
        assert (nc.cpu().numpy() == 44).all() and (v.sum(1) == 7).all()
        assert ((v > 0).sum(1) == 1).all()  # delta on the arg-max-prior child (Appendix C.1)


def test_search_cfg3_size_properties(mods, xo):
    """cfg 3 size (4,096 trees x 15 sims) and a cfg 4 slice (2,048 x 50): determinism, the
    n-8 visit law (B.4), and the oracle on a strided sample of trees."""
    eng, mcts = mods
    for n, n_sims, stride in ((4096, 15, 128), (2048, 50, 256)):
        bb = eng.BoardBatch(n)
        bb.playout(99, 9, capture_bias=40)          # diversify
        s = mcts.BatchedMCTS(n, n_sims)
        mv1, vis1, nc1 = [t.clone() for t in s.search(bb.board, bb.meta, mcts.HashEvaluator())]
        mv2, vis2, nc2 = s.search(bb.board, bb.meta, mcts.HashEvaluator())
        assert bool((vis1 == vis2).all()) and bool((mv1 == mv2).all()) and bool((nc1 == nc2).all())
        v, c = vis1.cpu().numpy(), nc1.cpu().numpy()
        live = c > 0
        assert live.all()                             # nothing ends within 9 plies
        assert (v.sum(1) == n_sims - 8).all()
        boards, meta = bb.boards_host(), bb.meta_host()
        pos = lambda q: None if q < 0 else (int(q) // 9, int(q) % 9)
        for g in range(0, n, stride):
            e = xo.Env().load(boards[g].reshape(10, 9), int(meta["player"][g]), int(meta["move_count"][g]),
                              None, pos(meta["red_king"][g]), pos(meta["black_king"][g]),
                              int(meta["no_capture"][g]))
            om, ov, _ = xo.mcts_search(e, n_sims)
            assert np.array_equal(mv1[g, :c[g]].cpu().numpy(), om) and np.array_equal(v[g, :c[g]], ov), g


def test_batched_opponent_mode_vs_oracle_game_loop(mods, xo):
    """Opponent mode of the batched game loop (self_play.py:190-198,211,234: red searches with
    `network`, black with `opponent_network`, only red's positions become samples, step rewards
    indexed by SAMPLE index) against the same loop on the oracle with the two hashed evaluators:
    move by move, outcome and shaped sample rewards."""
    eng, mcts = mods
    from chinesechessai_b200 import self_play
    for n, n_sims, temp, seed, first in ((20, 15, 1.0, 3, 0), (8, 30, 0.5, 4, 500)):
        sp = self_play.BatchedSelfPlay(mcts.HashEvaluator(False), n, n_sims, temperature=temp,
                                       opponent_network=mcts.HashEvaluator(True), seed=seed, first_game_id=first)
        sp.play()
        res = sp.materialise(red_only=True)
        P = sp.plies
        rm, rv, rn = (t[:P].cpu().numpy() for t in (sp.rec_moves, sp.rec_visits, sp.rec_n))
        played, rmove = sp.rec_played[:P].cpu().numpy(), sp.rec_move[:P].cpu().numpy()
        meta = sp.boards.meta_host()
        for g, (data, winner, reason) in enumerate(res):
            e = xo.Env()
            rewards, red_boards = [], []
            for p in range(70):
                red = e.current_player == 1
                om, ov, _ = xo.mcts_search(e, n_sims, flat=not red)     # red: hashed priors, black: flat
                if len(e.legal_moves_packed()) == 0 or len(om) == 0:
                    break
                assert played[p, g], (g, p)
                k = int(rn[p, g])
                assert np.array_equal(rm[p, g, :k], om) and np.array_equal(rv[p, g, :k], ov), (g, p)
                idx = xo.sample_move(ov, temp, seed, first + g, p)
                assert int(rmove[p, g]) == int(om[idx]), (g, p)
                if red:
                    red_boards.append(e.board.copy())
                rw, _, done = e.make_move(int(om[idx]))
                rewards.append(rw)
                if done:
                    break
            assert not played[len(rewards):, g].any() and played[:len(rewards), g].all(), g
            w = 0 if e.winner is None else e.winner
            assert winner == w and meta["reason"][g] == e.s.reason, g
            assert len(data) == len(red_boards)
            for i, (b, probs, total) in enumerate(data):
                assert np.array_equal(b, red_boards[i])
                # game_length = number of SAMPLES, immediate reward = step_rewards[sample index]
                assert repr(total) == repr(self_play.final_reward(w, 1, len(data)) + rewards[i] * 0.01), (g, i)


def test_batched_play_match_vs_oracle_loop(mods, xo):
    """chinesechessai_b200.compare_models.play_match (all games of a match as one device batch)
    against compare_models.py:35-92 restated on the oracle: same winners and move counts, hence
    the same result dict."""
    eng, mcts = mods
    from chinesechessai_b200.compare_models import MATCH_TEMPERATURE, play_match
    n, n_sims, seed = 10, 15, 21
    got = play_match(mcts.HashEvaluator(False), mcts.HashEvaluator(True), num_games=n, verbose=False,
                     num_simulations=n_sims, seed=seed)
    wins1 = wins2 = draws = total_moves = 0
    for g in range(n):
        e = xo.Env()
        for p in range(100):
            if len(e.legal_moves_packed()) == 0 or e.winner is not None:
                break
            om, ov, _ = xo.mcts_search(e, n_sims, flat=e.current_player != 1)
            if len(om) == 0:
                break
            e.make_move(int(om[xo.sample_move(ov, MATCH_TEMPERATURE, seed, g, p)]))
        total_moves += e.s.move_count
        wins1 += e.winner == 1
        wins2 += e.winner == -1
        draws += e.winner not in (1, -1)
    assert got["model1_wins"] == wins1 and got["model2_wins"] == wins2 and got["draws"] == draws
    assert got["avg_moves"] == total_moves / n
    assert got["model1_winrate"] == wins1 / n * 100 and got["draw_rate"] == draws / n * 100


def test_leaf_compaction_keeps_visit_counts(mods, xo):
    """BatchedMCTS with leaf compaction (only the games whose wave reached a network leaf go
    through the evaluator) gives the same trees as the full-batch search, on a ragged batch:
    finished games, inactive games and end-of-game positions where whole waves end on terminal
    leaves."""
    import torch
    eng, mcts = mods
    n, n_sims = 600, 50
    bb = eng.BoardBatch(n)
    bb.playout(7, 40, capture_bias=200)             # some games are already over
    late = eng.BoardBatch(n)
    late.playout(8, 67)                             # 3 plies before the cap: terminal children everywhere
    bb.board[n // 2:] = late.board[n // 2:]
    bb.meta[n // 2:] = late.meta[n // 2:]
    active = torch.ones(n, dtype=torch.uint8, device=bb.board.device)
    active[::7] = 0
    full = mcts.BatchedMCTS(n, n_sims)
    mv0, vis0, nc0 = [t.clone() for t in full.search(bb.board, bb.meta, mcts.HashEvaluator(), active)]
    meta = bb.meta_host()
    over = int((meta["winner"] != 2).sum())
    assert over >= 5                                # the batch really is ragged
    comp = mcts.BatchedMCTS(n, n_sims)
    # a bound of n (no compaction), a loose one and the tight one
    live = int(((meta["winner"] == 2) & (active.cpu().numpy() != 0)).sum())
    for bound in (n, n - 1, live):
        comp.rows_evaluated = 0
        mv, vis, nc = comp.search(bb.board, bb.meta, mcts.HashEvaluator(), active, rows=comp.bucket(bound))
        assert torch.equal(mv, mv0) and torch.equal(vis, vis0) and torch.equal(nc, nc0), bound
        assert comp.rows_evaluated == comp.bucket(bound) * 7
    assert comp.bucket(live) < n


def test_self_play_with_and_without_compaction(mods):
    """The batched game loop records the same games whether or not finished games are compacted
    out of the evaluator's batch; staggered openings make the batch ragged (games end at
    different plies)."""
    import torch
    eng, mcts = mods
    from chinesechessai_b200.self_play import BatchedSelfPlay
    runs = []
    for compact in (False, True):
        sp = BatchedSelfPlay(mcts.HashEvaluator(), 160, 15, temperature=1.0, seed=11, compact=compact,
                             use_graph=False)
        sp.restart(opening_seed=5, opening_plies=0)
        lib, b = sp.lib, sp.boards
        for k in range(4):                            # game g starts after (g // 40) * 16 random plies
            lo = k * 40
            res = torch.zeros((40, 40), dtype=torch.uint8, device=b.board.device)
            from chinesechessai_b200._lib import check
            check(lib.xq_playout(b.board[lo:].data_ptr(), b.meta[lo:].data_ptr(), b.pos_hist[lo:].data_ptr(),
                                 b.hist_cap, 5, lo, 16 * k, 0, res.data_ptr(), None, None, None, None, None,
                                 None, 40, torch.cuda.current_stream().cuda_stream))
        sp.play()
        runs.append((sp.plies, sp.rec_move[:sp.plies].clone(), sp.rec_visits[:sp.plies].clone(),
                     sp.rec_played[:sp.plies].clone(), sp.boards.board.clone(), sp.mcts.rows_evaluated))
    a, b = runs
    assert a[0] == b[0] and all(torch.equal(x, y) for x, y in zip(a[1:5], b[1:5]))
    assert b[5] < a[5]                               # fewer rows went through the evaluator


def test_bias_residual_relu_kernel(mods):
    """xq_bias_residual_relu_bf16 == relu(y + bias + x) computed in fp32 and rounded once."""
    import torch
    eng, _ = mods
    torch.manual_seed(3)
    for n in (1, 37, 4096):
        y = torch.randn(n, 128, 10, 9, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        x = torch.randn(n, 128, 10, 9, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        b = torch.randn(128, device="cuda").bfloat16()
        out = eng.bias_residual_relu(y, x, b)
        ref = torch.relu(y.float() + b.float()[None, :, None, None] + x.float()).bfloat16()
        assert out.is_contiguous(memory_format=torch.channels_last) and torch.equal(out, ref)


def test_folded_net_matches_module(mods):
    """The bf16 inference copy (BN folded, fused cuDNN calls, padded head, own residual epilogue)
    against the fp32 module: same arg-max move, logits within bf16 noise (numerics of the network
    are outside the bit-parity gate)."""
    import torch
    eng, mcts = mods
    from chinesechessai_b200.neural_network import ChessNet
    torch.manual_seed(0)
    net = ChessNet().cuda().eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(); m.running_var.uniform_(0.5, 2.0)
    bb = eng.BoardBatch(256)
    bb.playout(5, 12)
    planes = eng.encode_planes(bb.board, bb.meta[:, 0].view(torch.int8))
    with torch.no_grad():
        ref_p, ref_v = net(planes)
        f = mcts._FoldedNet(net, torch.bfloat16)
        assert f.fused and f.own_epilogue
        assert f.in_channels == 16  # zero-padded stem fed by the channels-last encode kernel
        pl = bb.meta[:, 0].view(torch.int8)
        p, v = f(eng.encode_planes_nhwc16(bb.board, pl))
        # the production entry: raw boards in, stem = fused encode + conv1 lookup kernel
        assert f.stem_table is not None
        p2, v2 = f.forward_boards(bb.board, pl)
        assert float((p2.float() - p.float()).abs().max()) < 0.05 and float((v2.float() - v.float()).abs().max()) < 0.02
        a = f._cr(f.stem, eng.encode_planes_nhwc16(bb.board, pl)).float()
        b = eng.stem_lookup(bb.board, pl, f.stem_table, f.stem_bias)
        assert b.shape == a.shape and b.is_contiguous(memory_format=torch.channels_last)
        d = (a - b.float()).abs()
        # same bf16 weights, float32 accumulation in a different order: at most one bf16 ulp apart
        assert float((d / a.abs().clamp(min=1.0)).max()) < 2 ** -7 and float((d > 0).float().mean()) < 0.05
        p, v = p2, v2
    assert p.shape == (256, 8192) and float(p[:, 8100:].abs().max()) == 0.0
    assert float((p[:, :8100].float() - ref_p).abs().max()) < 0.15
    assert float((v.float() - ref_v).abs().max()) < 0.05
    mv, nm = bb.legal_moves()
    pri_ref = eng.policy_priors(ref_p.contiguous(), mv, nm)
    pri = eng.policy_priors(p, mv, nm)
    assert float((pri - pri_ref).abs().max()) < 0.02
    assert float((pri.argmax(1) == pri_ref.argmax(1)).float().mean()) > 0.95


def test_folded_float32_copy_matches_module(mods):
    """The folded inference copy at float32 (BN folded, channels-last, fused cuDNN calls; what the
    TF32 evaluator runs): with TF32 off it equals the module up to the re-association of the BN
    fold; with TF32 on it stays within TF32 rounding and the search gives the n-8 visit law."""
    import torch
    eng, mcts = mods
    from chinesechessai_b200.neural_network import ChessNet
    torch.manual_seed(0)
    net = ChessNet().cuda().eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(); m.running_var.uniform_(0.5, 2.0)
    bb = eng.BoardBatch(256)
    bb.playout(5, 12)
    pl = bb.meta[:, 0].view(torch.int8).contiguous()
    mv, nm = bb.legal_moves()
    strict = mcts.NetEvaluator(net, torch.float32, tf32=False)
    assert not strict.folded
    pri_ref, val_ref = strict(bb.board, pl, mv, nm)
    folded = mcts.NetEvaluator(net, torch.float32, tf32=False, folded=True)
    pri, val = folded(bb.board, pl, mv, nm)
    assert folded._fast is not None and folded._fast.fused and not folded._fast.own_epilogue
    assert folded._fast.policy_fc.weight.dtype == torch.float32
    assert float((pri - pri_ref).abs().max()) < 2e-5 and float((val - val_ref).abs().max()) < 2e-4
    fast = mcts.NetEvaluator(net, torch.float32, tf32=True)
    assert fast.folded
    pri_t, val_t = fast(bb.board, pl, mv, nm)
    assert float((pri_t - pri_ref).abs().max()) < 5e-3 and float((val_t - val_ref).abs().max()) < 2e-2
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    fast(bb.board, pl, mv, nm)
    assert prev == (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    s = mcts.BatchedMCTS(32, 15)
    b32 = eng.BoardBatch(32)
    _, vis, nc = s.search(b32.board, b32.meta, fast)
    assert (nc.cpu().numpy() == 44).all() and (vis.cpu().numpy().sum(1) == 7).all()

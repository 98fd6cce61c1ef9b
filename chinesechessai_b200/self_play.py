"""Drop-in ``MCTS`` / ``self_play_game`` / ``parallel_self_play`` (reference: self_play.py).

* ``MCTS(network).search(env)`` — one tree on the GPU (xq_mcts_* kernels), the evaluator is
  whatever ``network.predict_batch`` the caller supplies, called exactly as the reference calls
  it (once per wave with one entry per queued simulation, self_play.py:143).
* ``self_play_game`` — the reference's game loop (self_play.py:178-312) on top of those two.
* ``parallel_self_play`` — the multi-process pool of the reference (self_play.py:368-469) becomes
  ONE batched device loop: all games advance together, one ``ChessNet.forward`` per MCTS wave for
  the whole batch, boards never leave HBM until the samples are materialised at the end.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BOARD_STRIDE, MAX_MOVES, META_DTYPE, check
from .chess_env import ChineseChess, format_end_reason
from .config import MAX_MOVES as MAX_PLIES, MCTS_SIMULATIONS
from .engine import BoardBatch, _ptr, _stream, pack_move, unpack_move
from .mcts import WAVE, BatchedMCTS, HashEvaluator, NetEvaluator

GRAPH_MAX_GAMES = 1024  # measured (scripts/single_game_latency.py): a ply is 4x
                        # faster at 1-64 games, 1.3x at 1,024, 1.05x at 4,096 — but the capture
                        # costs ~10 ms per BatchedSelfPlay instance, which a 70-ply batch only
                        # earns back below ~1,000 games

Move = Tuple[int, int, int, int]


class InterruptedWithResults(Exception):
    """Raised on Ctrl-C, carrying the games finished so far (self_play.py:12-16)."""

    def __init__(self, results: List[Tuple]) -> None:
        self.results = results
        super().__init__("Training interrupted by user")


def _env_state(env) -> Tuple[np.ndarray, np.ndarray]:
    """Root state the way MCTS._copy_env reads it (self_play.py:161-172); works for our
    ChineseChess and for the reference's own class."""
    b = np.zeros((1, BOARD_STRIDE), np.int8)
    b[0, :90] = np.asarray(env.board, dtype=np.int8).reshape(90)
    sq = lambda p: -1 if p is None else int(p[0]) * 9 + int(p[1])
    m = np.zeros(1, META_DTYPE)
    m["player"] = 1 if env.current_player == 1 else -1
    m["winner"] = _lib.WINNER_NONE if env.winner is None else int(env.winner)
    m["red_king"], m["black_king"] = sq(env.red_king_pos), sq(env.black_king_pos)
    m["move_count"], m["no_capture"] = int(env.move_count), int(env.no_capture_count)
    return b, m


class MCTS:
    def __init__(self, network, num_simulations=None):
        self.network = network
        self.default_simulations = num_simulations if num_simulations else MCTS_SIMULATIONS
        self._engines: Dict[int, BatchedMCTS] = {}

    def _engine(self, n_sims: int) -> BatchedMCTS:
        if n_sims not in self._engines:
            self._engines[n_sims] = BatchedMCTS(1, n_sims)
        return self._engines[n_sims]

    def search(self, env, num_simulations=None) -> Dict[Move, int]:
        """{move: visit_count} over the root's children in legal-move order, zeros included;
        {} if the root is terminal; ``env`` is not modified (self_play.py:89-154)."""
        n_sims = self.default_simulations if num_simulations is None else num_simulations
        if n_sims <= 0:
            return {}
        eng = self._engine(n_sims)
        d = eng.device
        b, m = _env_state(env)
        board = torch.from_numpy(b).to(d)
        meta = torch.from_numpy(m.view(np.uint8).reshape(1, 32)).to(d)
        eng.init(board, meta)
        priors = torch.zeros((1, MAX_MOVES), dtype=torch.float32, device=d)
        values = torch.zeros((1, WAVE), dtype=torch.float64, device=d)
        for start in range(0, n_sims, WAVE):
            eng.select(min(WAVE, n_sims - start))
            n_leaf = int(eng.leaf_n[0])
            if n_leaf == 0:
                continue  # every simulation of this wave ended in a terminal node
            mult = int(eng.leaf_mult[0])
            leaf_board = eng.leaf_board[0, :90].cpu().numpy().reshape(10, 9)
            player = int(eng.leaf_player[0])
            legal = [unpack_move(x) for x in eng.leaf_moves[0, :n_leaf].cpu().tolist()]
            # the reference queues the same leaf once per simulation of the wave (:138-143)
            results = self.network.predict_batch([(leaf_board.copy(), player, legal)] * mult)
            probs = results[0][0]
            p = np.zeros((1, MAX_MOVES), np.float32)
            p[0, :n_leaf] = [probs[mv] for mv in legal]
            v = np.zeros((1, WAVE), np.float64)
            v[0, :mult] = [float(r[1]) for r in results]
            priors.copy_(torch.from_numpy(p))
            values.copy_(torch.from_numpy(v))
            eng.backup(priors, values, values_per_game=WAVE)
        mv, vis, nc = eng.visits()
        k = int(nc[0])
        return {unpack_move(a): int(c) for a, c in zip(mv[0, :k].cpu().tolist(), vis[0, :k].cpu().tolist())}


def final_reward(winner: int, player: int, game_length: int) -> float:
    """Outcome part of the sample reward (self_play.py:268-298)."""
    if winner == 0:
        if game_length >= 60:
            return -0.15 if player == 1 else 0.05
        return -0.1 if player == 1 else 0.1
    if winner == player:
        bonus = 0.5 if game_length <= 30 else 0.3 if game_length <= 50 else 0.1 if game_length <= 70 else 0.0
        return 1.0 + bonus
    return -1.2 if game_length >= 60 else -1.0


def _shape_rewards(game_data, step_rewards, winner):
    """self_play.py:262-310 (step_rewards indexed by SAMPLE index, as the reference does)."""
    out = []
    n = len(game_data)
    for i, (board, move_probs, player) in enumerate(game_data):
        immediate = step_rewards[i] if i < len(step_rewards) else 0.0
        out.append((board, move_probs, final_reward(winner, player, n) + immediate * 0.01))
    return out


def self_play_game(network, temperature=1.0, render=False, num_simulations=None, opponent_network=None):
    """One game of self-play or network-vs-opponent (self_play.py:178-312).
    Returns ([(board int8[10,9], {move: prob}, reward)], winner, end_reason); the move is drawn
    with the global ``np.random`` generator exactly like the reference."""
    env = ChineseChess()
    mcts_red = MCTS(network, num_simulations=num_simulations)
    mcts_black = mcts_red if opponent_network is None else MCTS(opponent_network, num_simulations=num_simulations)
    game_data, step_rewards = [], []
    for _ in range(MAX_PLIES):
        board, player = env.get_state()
        if len(env.get_legal_moves()) == 0:
            break
        visit_counts = (mcts_red if player == 1 else mcts_black).search(env)
        if len(visit_counts) == 0:
            break
        moves = list(visit_counts.keys())
        counts = np.array(list(visit_counts.values()))
        if temperature < 0.01:
            move_probs = np.zeros(len(counts))
            move_probs[np.argmax(counts)] = 1
        else:
            counts = counts ** (1.0 / temperature)
            move_probs = counts / counts.sum()
        if player == 1 or opponent_network is None:
            game_data.append((board.copy(), {mv: p for mv, p in zip(moves, move_probs)}, player))
        move = moves[np.random.choice(len(moves), p=move_probs)]
        _, reward, done = env.make_move(move)
        step_rewards.append(reward)
        if render:
            env.render()
            print(f"走法: {move}, 即时奖励: {reward:.2f}, 访问次数: {visit_counts[move]}")
        if done:
            break
    winner = env.winner if env.winner else 0
    end_reason = env.end_reason if env.end_reason else "未知原因"
    return _shape_rewards(game_data, step_rewards, winner), winner, end_reason


# ---------------------------------------------------------------------------------------------
class BatchedSelfPlay:
    """All games of a shard played concurrently on one GPU.

    Per ply: one batched MCTS search (ceil(n_sims/8) select/forward/backup rounds for the whole
    batch), temperature sampling, one ``xq_step``.  Visit counts, sampled moves, per-ply boards,
    players and rewards are recorded in device tensors; ``materialise()`` turns them into the
    reference's sample tuples at the end."""

    def __init__(self, network, n_games: int, num_simulations: int, temperature: float = 1.0,
                 opponent_network=None, device: Optional[torch.device] = None,
                 net_dtype: torch.dtype = torch.float32, seed: Optional[int] = None,
                 first_game_id: int = 0, use_graph: Optional[bool] = None):
        self.n = int(n_games)
        self.n_sims = int(num_simulations)
        self.temperature = float(temperature)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = self.device
        self.lib = _lib.load()
        self.boards = BoardBatch(self.n, device=d, hist_cap=MAX_PLIES + 2)
        self.mcts = BatchedMCTS(self.n, self.n_sims, device=d)
        self.eval_red = network if callable(network) and not isinstance(network, torch.nn.Module) \
            else NetEvaluator(network, net_dtype)
        self.eval_black = None
        if opponent_network is not None:
            self.eval_black = opponent_network if callable(opponent_network) and not isinstance(
                opponent_network, torch.nn.Module) else NetEvaluator(opponent_network, net_dtype)
        # move sampling is counter-based: (seed, game id, ply) -> uniform, so a game's trajectory
        # does not depend on the batch or GPU it is played on
        self.seed = int(np.random.randint(0, 2**31 - 1)) if seed is None else int(seed)
        self.first_game_id = int(first_game_id)
        self.chosen = torch.zeros((self.n,), dtype=torch.int16, device=d)
        P = MAX_PLIES
        self.rec_board = torch.zeros((P, self.n, 90), dtype=torch.int8, device=d)
        self.rec_player = torch.zeros((P, self.n), dtype=torch.int8, device=d)
        self.rec_moves = torch.zeros((P, self.n, MAX_MOVES), dtype=torch.int16, device=d)
        self.rec_visits = torch.zeros((P, self.n, MAX_MOVES), dtype=torch.int32, device=d)
        self.rec_n = torch.zeros((P, self.n), dtype=torch.int16, device=d)
        self.rec_reward = torch.zeros((P, self.n), dtype=torch.float64, device=d)
        self.rec_move = torch.full((P, self.n), -1, dtype=torch.int16, device=d)
        self.rec_played = torch.zeros((P, self.n), dtype=torch.bool, device=d)
        self.plies = 0
        # game-loop state on the device: active[g] = 1 while game g is running, the move about to
        # be played, and a one-word "any game still running" flag (xq_selfplay_commit / _finish)
        self.active = torch.ones(self.n, dtype=torch.uint8, device=d)
        self.move = torch.full((self.n,), -1, dtype=torch.int16, device=d)
        self.any_active = torch.ones(1, dtype=torch.int32, device=d)
        # Small batches are launch-bound (about 60 launches per ply against ~1 ms of device
        # work up to a few hundred games): the whole search of a ply — init, every wave's
        # select / encode / forward / priors / backup, root visits — is captured once as a CUDA
        # graph and replayed.  Large batches are device-bound and stay eager.
        self.use_graph = (self.n <= GRAPH_MAX_GAMES and self.eval_black is None and
                          isinstance(self.eval_red, (NetEvaluator, HashEvaluator))) if use_graph is None \
            else bool(use_graph)
        self._graph = None
        self._graph_key = None
        self.graph_kernels = 0   # xq:: kernel launches captured in the graph
        self.graph_replays = 0   # (the library's launch counter only sees the capture)

    @property
    def done(self) -> torch.Tensor:
        return self.active == 0

    def _search(self):
        active = self.active
        if not self.use_graph or self.n == 0:
            return self._search_eager(active)
        key = (getattr(self.eval_red, "version", 0), getattr(self.eval_black, "version", 0))
        if self._graph is None or self._graph_key != key:
            self._search_eager(active)  # warm-up: library handles, workspaces, folded net
            torch.cuda.current_stream(self.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            c0 = self.lib.xq_launch_count()
            try:
                with torch.cuda.graph(graph):
                    self._search_eager(active)
            except Exception as exc:  # an evaluator that cannot be captured (host syncs, ...)
                import warnings
                warnings.warn(f"CUDA-graph capture of the search failed ({exc}); running eagerly")
                self.use_graph = False
                return self._search_eager(active)
            self._graph, self._graph_key = graph, key
            self.graph_kernels = int(self.lib.xq_launch_count() - c0)  # xq:: kernels per replay
        self._graph.replay()
        self.graph_replays += 1
        m = self.mcts
        return m.root_moves, m.root_visits, m.root_n

    def _search_eager(self, active: torch.Tensor):
        m, b = self.mcts, self.boards
        m.init(b.board, b.meta, active)
        if self.eval_black is None:
            for start in range(0, self.n_sims, WAVE):
                m.select(min(WAVE, self.n_sims - start))
                pri, val = self.eval_red(m.leaf_board, m.leaf_player, m.leaf_moves, m.leaf_n)
                m.backup(pri, val)
        else:  # red's tree uses `network`, black's uses the opponent (self_play.py:195-211)
            red = (b.meta[:, 0].view(torch.int8) == 1)
            for start in range(0, self.n_sims, WAVE):
                m.select(min(WAVE, self.n_sims - start))
                pr, vr = self.eval_red(m.leaf_board, m.leaf_player, m.leaf_moves, m.leaf_n)
                pb, vb = self.eval_black(m.leaf_board, m.leaf_player, m.leaf_moves, m.leaf_n)
                pri = torch.where(red[:, None], pr, pb).contiguous()
                val = torch.where(red, vr.double(), vb.double()).contiguous()
                m.backup(pri, val)
        return m.visits()

    @torch.no_grad()
    def play(self, max_plies: int = MAX_PLIES, check_done: bool = True) -> None:
        """Advance every unfinished game by up to ``max_plies`` further plies (the reference's
        cap of MAX_MOVES plies per game applies to the total).  ``check_done=False`` skips the
        per-ply "all games over?" host read (benchmarks that time a fixed number of plies)."""
        b, lib, n = self.boards, self.lib, self.n
        # "any game still running?" is read from pinned memory two plies late, so the host keeps
        # enqueueing ahead of the device instead of draining it every ply; the plies launched
        # past the end are no-ops (every game inactive) and are not counted
        lag = 2
        first = self.plies
        flag = torch.ones((MAX_PLIES,), dtype=torch.int32).pin_memory() if check_done else None
        events = {}
        for ply in range(self.plies, min(MAX_PLIES, self.plies + max_plies)):
            mv, vis, nc = self._search()
            with torch.cuda.device(self.device):
                st = _stream()
                # self_play.py:219-243 on device; idx = -1: game over / no legal move / empty search
                check(lib.xq_sample_moves(_ptr(vis), _ptr(nc), _ptr(self.active), self.temperature,
                                          self.seed, self.first_game_id, ply, _ptr(self.chosen), n, st))
                # :229-231 sample capture into this ply's rows + the move to play
                check(lib.xq_selfplay_commit(
                    _ptr(mv), _ptr(vis), _ptr(nc), _ptr(self.chosen), _ptr(b.board), _ptr(b.meta),
                    _ptr(self.rec_board[ply]), _ptr(self.rec_player[ply]), _ptr(self.rec_moves[ply]),
                    _ptr(self.rec_visits[ply]), _ptr(self.rec_n[ply]), _ptr(self.rec_played[ply]),
                    _ptr(self.rec_move[ply]), _ptr(self.move), _ptr(self.any_active), n, st))
                # :245 make_move; the step reward goes straight into this ply's row
                check(lib.xq_step(_ptr(b.board), _ptr(b.meta), _ptr(b.pos_hist), b.hist_cap,
                                  _ptr(self.move), _ptr(self.rec_reward[ply]), _ptr(b.flags), None, None,
                                  n, st))
                # :254-255 retire finished games
                check(lib.xq_selfplay_finish(_ptr(self.move), _ptr(b.flags), _ptr(self.active),
                                             _ptr(self.any_active), n, st))
            self.plies = ply + 1
            if check_done:
                flag[ply:ply + 1].copy_(self.any_active, non_blocking=True)
                events[ply] = torch.cuda.Event()
                events[ply].record()
                q = ply - lag
                if q >= first:
                    events.pop(q).synchronize()
                    if int(flag[q]) == 0:
                        self.plies = q + 1
                        break
        if check_done:  # the last `lag` plies were not looked at inside the loop
            for q in sorted(events):
                events[q].synchronize()
                if int(flag[q]) == 0:
                    self.plies = min(self.plies, q + 1)
                    break

    def stats(self) -> Dict[str, int]:
        plies = int(self.rec_played[:self.plies].sum())
        return {"plies": plies, "sims": plies * self.n_sims}

    def materialise(self, red_only: bool = False) -> List[Tuple[list, int, str]]:
        """-> [(game_data, winner, end_reason)] in the reference's format (self_play.py:312)."""
        P = self.plies
        rb = self.rec_board[:P].cpu().numpy()
        rp = self.rec_player[:P].cpu().numpy()
        rm = self.rec_moves[:P].cpu().numpy()
        rv = self.rec_visits[:P].cpu().numpy()
        rn = self.rec_n[:P].cpu().numpy()
        rr = self.rec_reward[:P].cpu().numpy()
        played = self.rec_played[:P].cpu().numpy()
        meta = self.boards.meta_host()
        out = []
        for g in range(self.n):
            game_data, step_rewards = [], []
            for p in range(P):
                if not played[p, g]:
                    break
                k = int(rn[p, g])
                counts = rv[p, g, :k].astype(np.int64)
                if self.temperature < 0.01:
                    probs = np.zeros(k)
                    probs[np.argmax(counts)] = 1
                else:
                    c = counts ** (1.0 / self.temperature)
                    probs = c / c.sum()
                player = int(rp[p, g])
                if player == 1 or not red_only:
                    game_data.append((rb[p, g].reshape(10, 9).copy(),
                                      {unpack_move(m): pr for m, pr in zip(rm[p, g, :k].tolist(), probs)},
                                      player))
                step_rewards.append(float(rr[p, g]))
            w = int(meta["winner"][g])
            winner = 0 if w == _lib.WINNER_NONE else w
            reason = format_end_reason(int(meta["reason"][g]), int(meta["player"][g]),
                                       int(meta["move_count"][g])) or "未知原因"
            out.append((_shape_rewards(game_data, step_rewards, winner), winner, reason))
        return out


def parallel_self_play(network, num_games, temperature=1.0, num_simulations=None, num_workers=4,
                       opponent_network=None):
    """Reference signature (self_play.py:368).  ``num_workers`` is accepted for compatibility;
    the games run as one device batch instead of a process pool.  Raises
    ``InterruptedWithResults`` on Ctrl-C with the games finished so far."""
    n_sims = num_simulations if num_simulations else MCTS_SIMULATIONS
    sp = BatchedSelfPlay(network, num_games, n_sims, temperature, opponent_network)
    try:
        sp.play()
    except KeyboardInterrupt:
        done = sp.boards.meta_host()["done"].astype(bool)
        res = sp.materialise(red_only=opponent_network is not None)
        raise InterruptedWithResults([r for r, d in zip(res, done) if d])
    return sp.materialise(red_only=opponent_network is not None)

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

import gc
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BOARD_STRIDE, MAX_MOVES, META_DTYPE, check
from .chess_env import ChineseChess, format_end_reason
from .config import MAX_MOVES as MAX_PLIES, MCTS_SIMULATIONS
from .engine import BoardBatch, _ptr, _stream, unpack_move
from .mcts import WAVE, BatchedMCTS, HashEvaluator, NetEvaluator

PROGRESS_EVERY_PLIES = 10  # parallel_self_play refreshes its progress line this often
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
    """``MCTS(network).search(env)`` (self_play.py:82-154) on the GPU tree kernels.

    ``network`` is anything with the reference's ``predict_batch``.  A stock ``ChessNet`` of this
    package (module on the GPU, ``predict_batch`` not overridden) takes the device path: the
    whole search — init, every wave's select / encode / forward / priors / backup, root visits —
    is captured once as a CUDA graph and replayed per call, with one 128-byte upload and one
    770-byte download around it (no per-wave host round trip).  Any other network object is
    called exactly as the reference calls it: once per wave, one list entry per queued
    simulation (:138-143).  Both give the same visit counts for the same network."""

    def __init__(self, network, num_simulations=None):
        self.network = network
        self.default_simulations = num_simulations if num_simulations else MCTS_SIMULATIONS
        self._engines: Dict[int, BatchedMCTS] = {}
        self._fast: Dict[int, "_DeviceSearch"] = {}

    def _engine(self, n_sims: int) -> BatchedMCTS:
        if n_sims not in self._engines:
            self._engines[n_sims] = BatchedMCTS(1, n_sims)
        return self._engines[n_sims]

    def _device_path(self) -> bool:
        from .neural_network import ChessNet
        net = self.network
        if not isinstance(net, ChessNet) or type(net).predict_batch is not ChessNet.predict_batch:
            return False
        if type(net).forward is not ChessNet.forward:
            return False
        p = next(net.parameters(), None)
        return p is not None and p.is_cuda and not net.training

    def search(self, env, num_simulations=None) -> Dict[Move, int]:
        """{move: visit_count} over the root's children in legal-move order, zeros included;
        {} if the root is terminal; ``env`` is not modified (self_play.py:89-154)."""
        n_sims = self.default_simulations if num_simulations is None else num_simulations
        if n_sims <= 0:
            return {}
        b, m = _env_state(env)
        if self._device_path():
            fs = self._fast.get(n_sims)
            if fs is None:
                fs = self._fast[n_sims] = _DeviceSearch(self.network, n_sims)
            return fs.search(b, m)
        eng = self._engine(n_sims)
        d = eng.device
        board = torch.from_numpy(b).to(d)
        meta = torch.from_numpy(m.view(np.uint8).reshape(1, 32)).to(d)
        eng.init(board, meta)
        priors = torch.zeros((1, MAX_MOVES), dtype=torch.float32, device=d)
        values = torch.zeros((1, WAVE), dtype=torch.float64, device=d)
        for start in range(0, n_sims, WAVE):
            eng.select(min(WAVE, n_sims - start))
            # one packed read of the wave's leaf: count, multiplicity, side to move, board, moves
            leaf = torch.cat([eng.leaf_n.view(torch.uint8), eng.leaf_mult.view(torch.uint8),
                              eng.leaf_player.view(torch.uint8), eng.leaf_board.view(torch.uint8).reshape(-1),
                              eng.leaf_moves.view(torch.uint8).reshape(-1)]).cpu().numpy()
            n_leaf, mult = int(leaf[0:2].view(np.int16)[0]), int(leaf[2:4].view(np.int16)[0])
            if n_leaf == 0:
                continue  # every simulation of this wave ended in a terminal node
            player = int(leaf[4:5].view(np.int8)[0])
            leaf_board = leaf[5:5 + 90].view(np.int8).reshape(10, 9)
            tup = _move_tuples()
            legal = [tup[x] for x in leaf[5 + BOARD_STRIDE:5 + BOARD_STRIDE + 2 * n_leaf].view(np.int16).tolist()]
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
        out = torch.cat([nc.view(torch.uint8), mv.view(torch.uint8).reshape(-1),
                         vis.view(torch.uint8).reshape(-1)]).cpu().numpy()
        k = int(out[0:2].view(np.int16)[0])
        tup = _move_tuples()
        moves = out[2:2 + 2 * MAX_MOVES].view(np.int16)[:k].tolist()
        counts = out[2 + 2 * MAX_MOVES:].view(np.int32)[:k].tolist()
        return {tup[a]: int(c) for a, c in zip(moves, counts)}


class _DeviceSearch:
    """One-tree search with the network on the device, replayed from a CUDA graph."""

    def __init__(self, network, n_sims: int):
        self.eng = BatchedMCTS(1, n_sims)
        self.ev = NetEvaluator(network, torch.float32)   # the module as given, like predict_batch
        d = self.d = self.eng.device
        self.state = torch.zeros(BOARD_STRIDE + 32, dtype=torch.uint8, device=d)
        self.board = self.state[:BOARD_STRIDE].view(torch.int8).reshape(1, BOARD_STRIDE)
        self.meta = self.state[BOARD_STRIDE:].reshape(1, 32)
        self.out = torch.zeros(2 + 2 * MAX_MOVES + 4 * MAX_MOVES, dtype=torch.uint8, device=d)
        self.h_state = torch.zeros(BOARD_STRIDE + 32, dtype=torch.uint8).pin_memory()
        self.h_out = torch.zeros_like(self.out, device="cpu").pin_memory()
        self._graph, self._key = None, None

    def _run(self) -> None:
        mv, vis, nc = self.eng.search(self.board, self.meta, self.ev)
        self.out.copy_(torch.cat([nc.view(torch.uint8), mv.view(torch.uint8).reshape(-1),
                                  vis.view(torch.uint8).reshape(-1)]))

    @torch.no_grad()
    def search(self, b: np.ndarray, m: np.ndarray) -> Dict[Move, int]:
        hs = self.h_state.numpy()
        hs[:BOARD_STRIDE] = b.view(np.uint8).reshape(-1)
        hs[BOARD_STRIDE:] = m.view(np.uint8).reshape(-1)
        self.state.copy_(self.h_state, non_blocking=True)
        key = self.ev.version
        if self._graph is None or self._key != key:
            self._run()                                  # warm-up: library handles, workspaces
            torch.cuda.current_stream(self.d).synchronize()
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    self._run()
                self._graph, self._key = g, key
            except Exception:                            # not capturable here: stay eager
                self._graph = None
                self._run()
        if self._graph is not None:
            self._graph.replay()
        self.h_out.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream(self.d).synchronize()
        out = self.h_out.numpy()
        k = int(out[0:2].view(np.int16)[0])
        tup = _move_tuples()
        moves = out[2:2 + 2 * MAX_MOVES].view(np.int16)[:k].tolist()
        counts = out[2 + 2 * MAX_MOVES:].view(np.int32)[:k].tolist()
        return {tup[a]: int(c) for a, c in zip(moves, counts)}


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
                 first_game_id: int = 0, use_graph: Optional[bool] = None, compact: bool = True):
        self.n = int(n_games)
        self.n_sims = int(num_simulations)
        self.temperature = float(temperature)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = self.device
        self.lib = _lib.load()
        self.boards = BoardBatch(self.n, device=d, hist_cap=MAX_PLIES + 2)
        self.mcts = BatchedMCTS(self.n, self.n_sims, device=d)
        for net in (network, opponent_network):  # the reference's workers force eval mode (:339,:346)
            if isinstance(net, torch.nn.Module):
                net.eval()
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
        # number of games still running, as last seen by the host (read a few plies late from
        # `any_active`); games only ever finish, so it bounds the live games of every later ply
        self.live_bound = self.n
        self.compact = bool(compact)
        self._flag = None      # pinned host copy of any_active, one word per ply
        self.finished = False  # set by play() once it has seen every game over
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

    def restart(self, opening_seed: Optional[int] = None, opening_plies: int = 4,
                capture_bias: int = 0) -> None:
        """Start a fresh batch in place: every game back to the initial position and, if
        ``opening_seed`` is given, ``opening_plies`` random plies (the shared counter-based pick
        rule, xq_playout) so that the games of the batch differ.  With few simulations the visit
        distribution of a fresh search is a delta on the arg-max-prior child, so games that start
        from the same position with the same network are all one trajectory (SURVEY §8d cfg 3)."""
        self.boards.reset()
        if opening_seed is not None and opening_plies > 0:
            self.boards.playout(int(opening_seed), int(opening_plies), first_game_id=self.first_game_id,
                                capture_bias=capture_bias)
        self.plies = 0
        self.finished = False
        self.live_bound = self.n
        self.active.fill_(1)
        self.any_active.fill_(self.n)
        self.rec_played.zero_()

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
            # leaf compaction: at most `live_bound` games are still running, so at most that many
            # rows can carry a network leaf (never inside a captured graph: its shapes are fixed)
            rows = None
            if self.compact and self._graph is None and not torch.cuda.is_current_stream_capturing():
                rows = m.bucket(self.live_bound)
            for start in range(0, self.n_sims, WAVE):
                m.select(min(WAVE, self.n_sims - start))
                m.evaluate_and_backup(self.eval_red, rows)
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
        flag = None
        if check_done:
            if self._flag is None:
                self._flag = torch.ones((MAX_PLIES,), dtype=torch.int32).pin_memory()
            flag = self._flag
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
                    self.live_bound = min(self.live_bound, int(flag[q]))
                    if int(flag[q]) == 0:
                        self.plies = q + 1
                        self.finished = True
                        break
        if check_done:  # the last `lag` plies were not looked at inside the loop
            for q in sorted(events):
                events[q].synchronize()
                self.live_bound = min(self.live_bound, int(flag[q]))
                if int(flag[q]) == 0:
                    self.plies = min(self.plies, q + 1)
                    self.finished = True
                    break

    def stats(self) -> Dict[str, int]:
        plies = int(self.rec_played[:self.plies].sum())
        return {"plies": plies, "sims": plies * self.n_sims}

    def outcome_counts(self) -> torch.Tensor:
        """Device tensor int64[5] = (red wins, black wins, draws, plies played, games): the
        statistics ``Trainer.collect_self_play_data`` returns (trainer.py:211-296), without a
        host read — summed over ranks with one all_reduce by the multi-GPU loop."""
        w = self.boards.meta[:, 1].view(torch.int8)
        return torch.stack([(w == 1).sum(), (w == -1).sum(), ((w == 0) | (w == _lib.WINNER_NONE)).sum(),
                            self.rec_played[:self.plies].sum(),
                            torch.tensor(self.n, device=self.device)]).to(torch.int64)

    def materialise(self, red_only: bool = False, games: Optional[Sequence[int]] = None
                    ) -> List[Tuple[list, int, str]]:
        """-> [(game_data, winner, end_reason)] in the reference's format (self_play.py:312);
        ``games`` restricts it to those games of the batch (in that order)."""
        P = self.plies
        meta = self.boards.meta_host()
        if games is None:
            sel = lambda t: t[:P].cpu().numpy()
        else:
            idx = torch.as_tensor(list(games), dtype=torch.int64, device=self.device)
            sel = lambda t: t[:P].index_select(1, idx).cpu().numpy()
            meta = meta[np.asarray(list(games), dtype=np.int64)]
        return materialise_arrays(
            sel(self.rec_board), sel(self.rec_player), sel(self.rec_moves), sel(self.rec_visits),
            sel(self.rec_n), sel(self.rec_reward), sel(self.rec_played), meta["winner"],
            meta["reason"], meta["player"], meta["move_count"], self.temperature, red_only)


_MOVE_TUPLES: List[Move] = []


def _move_tuples() -> List[Move]:
    if not _MOVE_TUPLES:
        _MOVE_TUPLES.extend(unpack_move(m) for m in range(_lib.POLICY))
    return _MOVE_TUPLES


class _LazyMoveProbs(dict):
    """The ``{move: prob}`` dict of one sample (self_play.py:236), filled on first use.

    ``Trainer`` stores these dicts in its replay buffer and never reads them (trainer.py:311-321
    trains on boards and rewards only), so building 40-entry dicts of numpy scalars for every
    sample of a large batch would cost more host time than the games took on the GPU.  The
    packed moves and visit counts of the ply are kept instead; any read access, comparison, copy
    or pickle fills the dict with exactly the keys, order and float64 values the eager path
    produced."""
    __slots__ = ("_src",)

    def _fill(self) -> None:
        # _src = (moves int16[P,128], visits int32[P,128], ply, n_moves, temperature) or unset
        src = getattr(self, "_src", None)
        if src is None:
            return
        self._src = None
        moves, visits, p, k, temperature = src
        moves_row = moves[p, :k]
        counts = visits[p, :k].astype(np.int64)
        if temperature < 0.01:                                   # self_play.py:225-228
            probs = np.zeros(len(counts))
            probs[np.argmax(counts)] = 1
        else:                                                    # :230-231
            c = counts ** (1.0 / temperature)
            probs = c / c.sum()
        tup = _move_tuples()
        dict.update(self, zip((tup[m] for m in moves_row.tolist()), probs))

    def _filled(name):  # noqa: N805 — builds the forwarding methods below
        base = getattr(dict, name)

        def method(self, *a, **k):
            self._fill()
            return base(self, *a, **k)
        method.__name__ = name
        return method

    for _n in ("__getitem__", "__iter__", "__len__", "__contains__", "__eq__", "__ne__", "__repr__",
               "__reversed__", "__or__", "__ror__", "__setitem__", "__delitem__", "__ior__",
               "get", "items", "keys", "values", "copy", "pop", "popitem", "setdefault", "update", "clear"):
        locals()[_n] = _filled(_n)
    del _n, _filled

    def __bool__(self) -> bool:
        return self.__len__() > 0

    def __reduce__(self):  # pickles (data/best_games.pkl, trainer.py:487-497) as a plain dict
        self._fill()
        return (dict, (dict(self),))


def materialise_arrays(rb, rp, rm, rv, rn, rr, played, winner_m, reason_m, player_m, move_count_m,
                       temperature: float, red_only: bool = False, lazy: bool = True
                       ) -> List[Tuple[list, int, str]]:
    """Recorded plies of a batch (arrays [P, n, ...]) -> the reference's
    ``[(game_data, winner, end_reason)]`` (self_play.py:203-312), vectorised over all samples:
    reward shaping in float64 with the reference's constants and operation order, boards as
    views of one array, move-prob dicts lazily (``lazy=False`` builds them eagerly)."""
    P, n = played.shape
    out: List[Tuple[list, int, str]] = []
    if n == 0:
        return out
    played = played.astype(bool)
    # a game's plies are a prefix (a game never resumes): count instead of scanning for the gap
    n_plies = np.where(played.all(0), P, np.argmin(played, axis=0)) if P else np.zeros(n, np.int64)
    prefix = np.arange(P)[:, None] < n_plies[None, :]
    keep = prefix & ((rp == 1) if red_only else True)                        # :234
    n_samples = keep.sum(0)
    winner = np.where(winner_m == _lib.WINNER_NONE, 0, winner_m).astype(np.int64)   # :259
    # final reward by (winner, player, game_length = samples of the game) — :264-298
    L = n_samples[None, :]
    pl = rp.astype(np.int64)
    w = winner[None, :]
    draw = np.where(L >= 60, np.where(pl == 1, -0.15, 0.05), np.where(pl == 1, -0.1, 0.1))
    win = 1.0 + np.where(L <= 30, 0.5, np.where(L <= 50, 0.3, np.where(L <= 70, 0.1, 0.0)))
    lose = np.where(L >= 60, -1.2, -1.0)
    fin = np.where(w == 0, draw, np.where(w == pl, win, lose))
    # step_rewards is indexed by SAMPLE index (:303-304): the i-th kept sample of a game takes
    # the reward of the game's i-th ply
    sidx = np.cumsum(keep, axis=0) - 1
    imm = np.take_along_axis(rr, np.clip(sidx, 0, max(P - 1, 0)), axis=0) if P else rr
    imm = np.where(sidx < n_plies[None, :], imm, 0.0)
    total = fin + imm * 0.01                                                  # :308
    boards = np.ascontiguousarray(rb.transpose(1, 0, 2)).reshape(n, P, 10, 9)  # [n, P, 10, 9]
    total_t, keep_t = np.ascontiguousarray(total.T), np.ascontiguousarray(keep.T)
    rm_t, rv_t, rn_t = rm.transpose(1, 0, 2), rv.transpose(1, 0, 2), rn.T
    # millions of small containers are created below; the cyclic collector would rescan them
    # again and again (measured: 1.5 s with it, 0.24 s without, for 4,096 games x 70 plies)
    gc_was_on = gc.isenabled()
    gc.disable()
    try:
        _build_games(out, n, keep_t, total_t, boards, rm_t, rv_t, rn_t, temperature, lazy, winner,
                     reason_m, player_m, move_count_m)
    finally:
        if gc_was_on:
            gc.enable()
    return out


def _build_games(out, n, keep_t, total_t, boards, rm_t, rv_t, rn_t, temperature, lazy, winner,
                 reason_m, player_m, move_count_m) -> None:
    for g in range(n):
        ps = np.flatnonzero(keep_t[g]).tolist()
        rew = total_t[g].tolist()
        bg, mg, vg, kg = boards[g], rm_t[g], rv_t[g], rn_t[g].tolist()
        game_data = []
        for p in ps:
            probs = _LazyMoveProbs()
            probs._src = (mg, vg, p, kg[p], temperature)
            if not lazy:
                probs._fill()
            game_data.append((bg[p], probs, rew[p]))
        reason = format_end_reason(int(reason_m[g]), int(player_m[g]), int(move_count_m[g])) or "未知原因"
        out.append((game_data, int(winner[g]), reason))


def _progress(done: int, total: int, valid: int) -> None:
    """The reference's progress line (self_play.py:411-431): a 50-column bar, games finished and
    the share of them that produced data."""
    frac = done / total if total else 1.0
    bar = "=" * int(50 * frac) + " " * (50 - int(50 * frac))
    rate = valid / done * 100 if done else 0
    tail = f" | 有效:{valid} ({rate:.0f}%)" if done else ""
    print(f"\r   进度: [{bar}] {done}/{total} ({frac * 100:.1f}%){tail}", end="", flush=True)


def parallel_self_play(network, num_games, temperature=1.0, num_simulations=None, num_workers=4,
                       opponent_network=None):
    """Reference signature (self_play.py:368).  ``num_workers`` is accepted for compatibility;
    the games run as one device batch instead of a process pool.  Like the reference's workers
    (:339,:346) the networks are put in eval mode.  Progress is printed in the reference's format
    as games finish; Ctrl-C raises ``InterruptedWithResults`` carrying the games that had
    finished (:433-452)."""
    n_sims = num_simulations if num_simulations else MCTS_SIMULATIONS
    for net in (network, opponent_network):
        if isinstance(net, torch.nn.Module):
            net.eval()
    red_only = opponent_network is not None
    # The search runs the network as given — float32, the reference's precision — unless the
    # caller opts into a folded inference copy: "bf16" (about 10x the simulations per second) or
    # "tf32" (float32 storage, TF32 tensor cores); numerics of the forward are outside the parity
    # boundary, SURVEY B.5.
    prec = os.environ.get("XQ_SELFPLAY_DTYPE", "fp32")
    if prec not in ("fp32", "bf16", "tf32"):
        raise ValueError(f"XQ_SELFPLAY_DTYPE={prec!r}: expected fp32, tf32 or bf16")
    dtype = torch.bfloat16 if prec == "bf16" else torch.float32
    if prec == "tf32":
        network.eval()
        players = [NetEvaluator(network, torch.float32, tf32=True)]
        if opponent_network is not None:
            opponent_network.eval()
            players.append(NetEvaluator(opponent_network, torch.float32, tf32=True))
        sp = BatchedSelfPlay(players[0], num_games, n_sims, temperature,
                             players[1] if len(players) > 1 else None)
    else:
        sp = BatchedSelfPlay(network, num_games, n_sims, temperature, opponent_network, net_dtype=dtype)
    _progress(0, num_games, 0)
    try:
        # a few plies per slice so that the progress line moves and Ctrl-C is honoured promptly
        while sp.plies < MAX_PLIES and not sp.finished:
            sp.play(max_plies=PROGRESS_EVERY_PLIES)
            done = int((sp.active == 0).sum())
            _progress(done, num_games, done)
    except KeyboardInterrupt:
        print("\n\n⚠️  检测到Ctrl+C，正在停止对弈...", flush=True)
        torch.cuda.synchronize(sp.device)
        done = (sp.active == 0).cpu().numpy().astype(bool)
        res = sp.materialise(red_only=red_only)
        results = [r for r, d in zip(res, done) if d]
        print("✓ 已停止对弈", flush=True)
        print(f"提示: 已完成 {len(results)} 局对弈，数据将被保存", flush=True)
        raise InterruptedWithResults(results)
    _progress(num_games, num_games, num_games)
    print()
    return sp.materialise(red_only=red_only)


def test_self_play():
    """``python main.py test`` calls this (main.py:182): one rendered game with few simulations."""
    from .neural_network import ChessNet
    print("测试自我对弈系统...")
    network = ChessNet().to(torch.device("cuda"))
    network.eval()
    game_data, winner, end_reason = self_play_game(network, render=True, num_simulations=10)
    print(f"\n对局结束！ 总步数: {len(game_data)}  胜者: {['和局', '红方', '黑方'][winner]}  "
          f"结束原因: {end_reason}  训练样本: {len(game_data)}")

#!/usr/bin/env python
"""bench.py — board-steps/s of the Xiangqi rules hot path on B200 (BASELINE.json metric).

Workload (N=1): BASELINE.json configs[1] — batched random playouts, 65,536 boards x <=70 plies
from the initial position, uniform legal move per ply by the shared counter-based pick rule
(philox4x32-10 keyed by seed, counter (game_id, ply)); one "step" = one such batch through
the fused playout kernel (get_legal_moves + make_move per ply, every terminal rule).  N>1:
the same batch per GPU (weak scaling), game ids offset per rank, no data-path collective.

    python bench.py --gpus N --steps K --warmup W          # our arm (CUDA, no CPU fallback)
    python bench.py --impl reference ...                   # the reference's CPU implementation

Prints ONE JSON line.  ``value`` is device-resident throughput, ``e2e`` goes through the
host-buffer C-ABI call (H2D + kernel + D2H inside the timed region).  The line verifies itself:
the results of timed step 0 are compared, game by game, with the CPU oracle's (and, for a sample,
with the unmodified Python reference's) before anything is printed; a mismatch fails the run.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BOARDS = 65536
PLIES = 70
SEED = 0x5EED
# Algorithmic HBM bytes per board-step (DESIGN.md §3 / SURVEY.md §8d):
#   step-per-launch mode: state in 128 + state out 128 + legal list 81 + move 4 + history 8
BYTES_PER_STEP_LAUNCH_MODE = 352
#   fused playout: (board+meta in 128, out 128, result 40) / 70 plies + one 8-byte history append
BYTES_PER_STEP_FUSED = (128 + 128 + 40) / 70.0 + 8.0
FLOP_PER_LEAF_EVAL = 263_209_216          # ChessNet.forward, SURVEY.md §8d
MCTS_GAMES, MCTS_SIMS, MCTS_OPENING_PLIES = 4096, 15, 4
METRIC = "board-steps/sec (legal movegen+step)"
UNIT = "board-steps/s"
RESULT_FIELDS = ("plies", "winner", "reason", "max_legal", "digest", "final_hash")


def workload_config(n: int, world: int) -> dict:
    """The `config` object — the SAME in both arms (ours and --impl reference)."""
    return {"workload": f"cfg2: batched random playouts, {n} boards x <= {PLIES} plies per GPU from the "
                        "initial position (get_legal_moves + make_move per ply, every terminal rule)",
            "boards_per_gpu": n, "max_plies": PLIES, "seed": SEED,
            "pick": "philox4x32-10(seed,(game,ply)) mod n_legal",
            "parallelism": f"games sharded over {world} GPU(s), no collective on the path"}


def kernel_counts() -> dict:
    """ncu-derived constants of the fused playout kernels (warp-instructions per board-step, DRAM
    bytes per 65,536-board launch), regenerated from the captures under profiles/ by
    profiles/kernel_counts.py — not hard-coded here."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_counts.json")))
    except Exception:
        return {}


def read_peaks() -> dict:
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_burst": float(p["bf16_tflops"]),
                "bf16_sustained": float(p["bf16_tflops_sustained"]), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_burst": 1600.0, "bf16_sustained": 1400.0,
                "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arms: the C port of the reference path (oracle/) and the unmodified Python reference
def cpu_port(n_games: int, threads: int, seed: int = SEED, first: int = 0):
    """Oracle (C port of the reference path) on the host cores -> (plies/s, plies, s, results)."""
    from oracle import xq_oracle as xo
    xo.build()
    t0 = time.perf_counter()
    total, res = xo.playout_many(n_games, seed, first, PLIES, 0, n_threads=threads)
    dt = time.perf_counter() - t0
    return total / dt, total, dt, res


def py_reference(seconds: float, seed: int = SEED, first: int = 0):
    """The UNMODIFIED Python reference (chess_env.ChineseChess from baseline/_ref) on all host
    cores with multiprocessing, a bounded sample of the workload's game ids starting at `first`.
    None if the reference checkout is not available on this box."""
    from baseline import reference as R
    if R.locate() is None:
        return None
    from baseline import reference_playout as rp
    procs = os.cpu_count() or 1
    # fresh interpreters with CUDA hidden (this process may already hold a CUDA context)
    cal = rp.run_subprocess(first, procs, seed, PLIES)                         # one game per core
    per_round = max(cal["seconds"], 1e-3)
    rounds = int(max(1, min(8, round(seconds / per_round) - 1)))
    run = rp.run_subprocess(first + procs, procs * rounds, seed, PLIES)
    import numpy as np
    results = np.concatenate([cal["results"], run["results"]])
    plies, dt = cal["plies"] + run["plies"], cal["seconds"] + run["seconds"]
    return {"value": plies / dt, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": f"game ids {first}..{first + len(results) - 1} of the workload ({len(results)} games, "
                      f"{plies} plies in {dt:.1f} s), unmodified chess_env.py, multiprocessing x {procs}",
            "plies": plies, "seconds": dt, "first": first, "results": results,
            "path": R.locate()}


def same_results(a, b, fields=RESULT_FIELDS):
    """Field-by-field equality of two playout result arrays (reward sums as float64 bit patterns)."""
    import numpy as np
    bad = [f for f in fields if not np.array_equal(a[f], b[f])]
    if not np.array_equal(a["reward_sum"].view(np.uint64), b["reward_sum"].view(np.uint64)):
        bad.append("reward_sum")
    return bad


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The
    headline value is the UNMODIFIED Python reference (multiprocessing on all cores) when its
    checkout travelled to this box (baseline/_ref); the C port of the same algorithm
    (oracle/xq_oracle.c, pthreads) is timed beside it and is the headline when the Python
    reference is absent.  Each step is a bounded sample of the workload's games."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    # C port: K steps, each a sample sized for ~4 s
    rate, _, _, _ = cpu_port(1024, threads)                                   # calibrate
    games = int(max(256, min(args.boards, rate * 4.0 / 69.0)))
    for w in range(min(args.warmup, 2)):
        cpu_port(min(games, 512), threads, SEED + 1000 + w)
    p_plies, p_t = 0, 0.0
    for k in range(args.steps):
        _, plies, dt, _ = cpu_port(games, threads, SEED + k)
        p_plies += plies
        p_t += dt
    port = {"value": p_plies / p_t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{games} of {args.boards} games per step x {args.steps} steps, {p_plies} plies in "
                      f"{p_t:.1f} s (oracle/xq_oracle.c, pthreads)"}
    # Python reference: ~25 s in total, split over the steps
    ref = None
    if not args.no_python:
        ref = py_reference(max(8.0, min(30.0, 3.0 * args.steps)))
    verified = None
    if ref is not None:
        _, _, _, pres = cpu_port(len(ref["results"]), threads, SEED, ref["first"])
        bad = same_results(ref["results"], pres)
        verified = {"python_reference_equals_port": not bad, "games": len(ref["results"]), "fields_differing": bad}
        if bad:
            raise SystemExit(f"bench.py: the Python reference and the C port disagree on {bad}")
    head = ref if ref is not None else port
    cb = {k: head[k] for k in ("value", "unit", "cores", "kind", "sample")}
    cb["port"] = port
    if ref is None:
        cb["reference_unavailable"] = "no reference checkout on this box (baseline/_ref, XQ_REFERENCE)"
    value = head["value"]
    ms = 1e3 * (ref["seconds"] if ref is not None else p_t) / args.steps
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8/f64",
        "data": "synthetic", "config": workload_config(args.boards, world),
        "cpu_baseline": cb, "verified": verified,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------
def measure_step_per_launch(torch, BoardBatch, n, first_id, steps, flush):
    """API-faithful mode: ONE launch per ply — xq_step_pick applies the picked move of the list the
    previous launch left behind and writes the next list (352 algorithmic bytes/board-step); the
    70-launch loop is captured once as a CUDA graph and replayed."""
    bb = BoardBatch(n, hist_cap=PLIES + 2)
    bb.reset()
    bb.legal_moves()

    def loop(seed):
        for ply in range(PLIES):
            bb.step_pick(seed, ply, first_game_id=first_id)

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        loop(SEED + 999)                      # warm-up outside capture
    torch.cuda.synchronize()
    graphs = {}
    for k in range(steps):                    # the seed is a launch argument: one graph per step
        g = torch.cuda.CUDAGraph()
        bb.reset()
        bb.legal_moves()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            loop(SEED + k)
        graphs[k] = g
    ms, plies = 0.0, 0
    for k in range(steps):
        bb.reset()
        bb.legal_moves()
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graphs[k].replay()
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
        plies += int(bb.meta_host()["move_count"].sum())
    final = bb.meta_host()["move_count"].copy(), bb.boards_host()
    return plies / (ms * 1e-3), ms / steps, PLIES, final


def measure_tree_only(torch, dev, games, sims):
    """Search kernels alone (select / expand / backup + the hashed stand-in evaluator): sims/s of
    the tree machinery without the network, and the same literal algorithm (self_play.py:89-154,
    C port, hashed evaluator) on the host cores for a bounded sample."""
    from concurrent.futures import ThreadPoolExecutor
    from chinesechessai_b200.engine import BoardBatch
    from chinesechessai_b200.mcts import BatchedMCTS, HashEvaluator
    from oracle import xq_oracle as xo
    bb = BoardBatch(games, device=dev)
    bb.playout(SEED, MCTS_OPENING_PLIES)
    m = BatchedMCTS(games, sims, device=dev)
    ev = HashEvaluator()
    m.search(bb.board, bb.meta, ev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        m.search(bb.board, bb.meta, ev)
    b.record()
    torch.cuda.synchronize()
    gpu = games * sims * reps / (a.elapsed_time(b) * 1e-3)
    threads = os.cpu_count() or 1
    boards, meta = bb.boards_host(), bb.meta_host()
    pos = lambda q: None if q < 0 else (int(q) // 9, int(q) % 9)
    n_cpu = min(games, 64 * threads)

    def one(g):
        e = xo.Env().load(boards[g].reshape(10, 9), int(meta["player"][g]), int(meta["move_count"][g]),
                          None, pos(meta["red_king"][g]), pos(meta["black_king"][g]),
                          int(meta["no_capture"][g]))
        xo.mcts_search(e, sims)
    xo.mcts_search(xo.Env(), sims)
    t0 = time.perf_counter()
    done = 0
    with ThreadPoolExecutor(threads) as ex:
        while time.perf_counter() - t0 < 3.0:       # bounded sample: ~3 s of host work
            list(ex.map(one, range(n_cpu)))
            done += n_cpu
    dt = time.perf_counter() - t0
    return {"gpu_sims_per_s": gpu,
            "cpu_baseline": {"value": done * sims / dt, "unit": "sims/s", "cores": threads, "kind": "port",
                             "sample": f"{done} searches x {sims} sims, literal replay per simulation "
                                       f"(oracle/xq_oracle.c via ctypes threads), {dt:.1f} s"}}


PRECISIONS = ("bf16", "tf32", "fp32")


def make_evaluator(torch, net, precision):
    """bf16: BN-folded channels-last inference copy (own stem / residual-epilogue kernels);
    tf32: float32 storage, BN-folded channels-last copy, TF32 tensor cores for cuDNN and cuBLAS;
    fp32: the module as given, strict IEEE float32 everywhere (TF32 off) — the reference's
    arithmetic (neural_network.py:47-71)."""
    from chinesechessai_b200.mcts import NetEvaluator
    if precision == "bf16":
        return NetEvaluator(net, torch.bfloat16)
    return NetEvaluator(net, torch.float32, tf32=(precision == "tf32"))   # fp32 leg: TF32 forced OFF


def precision_agreement(torch, dev, net, n_pos=4096):
    """the forward is outside bit-parity; report arg-max agreement and max |d prior| /
    |d value| of each precision against strict fp32 on the same leaf positions."""
    from chinesechessai_b200.engine import BoardBatch
    bb = BoardBatch(n_pos, device=dev)
    bb.playout(SEED + 7, 12)                         # 12 random plies: varied middle-game-ish positions
    moves, n_moves = bb.legal_moves()
    player = bb.meta[:, 0].view(torch.int8).contiguous()
    outs = {}
    for prec in PRECISIONS:
        pri, val = make_evaluator(torch, net, prec)(bb.board, player, moves, n_moves)
        outs[prec] = (pri.float().clone(), val.float().clone())
    ref_p, ref_v = outs["fp32"]
    live = n_moves > 0
    res = {}
    for prec in ("bf16", "tf32"):
        p, v = outs[prec]
        res[prec] = {"argmax_agreement": float((p.argmax(1) == ref_p.argmax(1))[live].float().mean()),
                     "max_abs_dprior": float((p - ref_p).abs().max()),
                     "max_abs_dvalue": float((v - ref_v).abs().max())}
    res["positions"] = int(live.sum())
    res["reference"] = "strict fp32 (TF32 off) through the same kernels for encode and prior softmax"
    return res


def timed_selfplay(torch, dev, sp, min_seconds, min_plies, dist=None):
    """Advance `sp` ply by ply (no host read inside) until >= min_seconds of device time and
    >= min_plies plies have passed; new batches are started as games run out of plies.
    Returns (ms, plies played summed over games)."""
    a = torch.cuda.Event(enable_timing=True)
    played = torch.zeros((), dtype=torch.int64, device=dev)
    ms, plies_done, chunk = 0.0, 0, max(2, min_plies // 2)
    while ms < min_seconds * 1e3 or plies_done < min_plies:
        if sp.plies + chunk > PLIES - MCTS_OPENING_PLIES:   # every game is at the 70-ply cap: fresh openings
            sp.restart(SEED + plies_done, MCTS_OPENING_PLIES)
        p0 = sp.plies
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        sp.play(chunk, check_done=False)
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
        played += sp.rec_played[p0:sp.plies].sum()
        plies_done += sp.plies - p0
    return ms, int(played), plies_done


def measure_mcts(torch, dev, precision, games=MCTS_GAMES, sims=MCTS_SIMS, label="cfg3", min_seconds=2.0,
                 peaks=None):
    """cfg 3: 4,096 concurrent self-play games, 15 sims/move (2 waves of 8+7), random-init ChessNet
    (torch.manual_seed(0)), temperature 1.0; the batch is diversified by 4 random opening plies
    (otherwise all games are one trajectory).  cfg 4 (per GPU): 16,384 games, 50
    sims/move (7 waves).  Timed for >= min_seconds on the device with the clock sampler running."""
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.self_play import BatchedSelfPlay
    torch.manual_seed(0)
    net = ChessNet().to(dev).eval()
    sp = BatchedSelfPlay(make_evaluator(torch, net, precision), games, sims, temperature=1.0, device=dev, seed=0)
    sp.restart(SEED, MCTS_OPENING_PLIES)
    sp.play(2, check_done=False)                         # warm-up plies (cuDNN heuristics, workspaces)
    torch.cuda.synchronize()
    clocks = ClockSampler(dev.index or 0).start()
    ms, played, plies_timed = timed_selfplay(torch, dev, sp, min_seconds, 4)
    clk = clocks.stop()
    waves = (sims + 7) // 8
    evals = played * waves
    tf = evals / (ms * 1e-3) * FLOP_PER_LEAF_EVAL / 1e12
    peaks = peaks or read_peaks()
    out = {"metric": "MCTS sims/sec", "value": played * sims / (ms * 1e-3), "unit": "sims/s",
           "precision": precision,
           "config": {"workload": f"{label}: {games} concurrent games, {sims} sims/move, random-init ChessNet, "
                                  f"T=1.0, {MCTS_OPENING_PLIES} random opening plies",
                      "plies_timed": plies_timed, "seconds_timed": ms * 1e-3,
                      "cuda_graph": bool(sp.use_graph)},
           "ms_per_ply": ms / plies_timed, "unique_leaf_evals_per_s": evals / (ms * 1e-3),
           "leaf_rows_evaluated_per_s": sp.mcts.rows_evaluated / (ms * 1e-3) if getattr(sp.mcts, "rows_evaluated", 0) else None,
           "clocks": clk,
           "roofline": {"bound": "tensor", "achieved": tf, "unit": "TFLOP/s",
                        "peak": peaks["bf16_sustained"], "frac": tf / peaks["bf16_sustained"],
                        "peak_kind": "bf16 dense, SUSTAINED (cuBLAS back to back for 4 s, MEASURED_PEAKS.json): "
                                     "the leg is timed for >= 2 s; against the burst peak "
                                     f"({peaks['bf16_burst']:.0f}) the fraction is {tf / peaks['bf16_burst']:.3f}",
                        "flop_per_leaf_eval": FLOP_PER_LEAF_EVAL, "traffic": None}}
    if precision != "bf16":
        out["roofline"]["note"] = ("the denominator is the bf16 peak (no fp32/TF32 peak is measured on this pool); "
                                   "nominal dense TF32 is half of bf16, fp32 SIMT ~80 TFLOP/s")
    return out


def measure_mcts_ragged(torch, dev, games=MCTS_GAMES, sims=MCTS_SIMS, strata=8, step=8):
    """Leaf compaction on a RAGGED batch: game g starts after (g * strata // games) * step random
    opening plies, so the 70-ply cap retires one eighth of the batch every 8 plies (7/8 of the
    games end before the batch's last ply; average live fraction 56 %).  The reference never sends
    a finished game or a terminal leaf to the network (self_play.py:126-139); without compaction a
    batch runs the forward on all rows every wave.  Same games, same visit counts either way
    (tests/test_mcts_gpu.py); bf16 network, complete batch timed with CUDA events, host loop with
    its two-ply-late 'games still running' read included."""
    from chinesechessai_b200._lib import check
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.self_play import BatchedSelfPlay
    torch.manual_seed(0)
    net = ChessNet().to(dev).eval()
    ev = make_evaluator(torch, net, "bf16")
    out = {"workload": f"{games} games x {sims} sims/move, openings of 0,{step},..,{step * (strata - 1)} random plies "
                       f"by stratum ({strata} strata): games retire at batch plies 70,{70 - step},..; bf16"}
    for compact in (False, True):
        sp = BatchedSelfPlay(ev, games, sims, temperature=1.0, device=dev, seed=0, compact=compact, use_graph=False)

        def start():
            sp.restart(None)
            b = sp.boards
            per = games // strata
            for k in range(1, strata):
                lo = k * per
                cnt = games - lo if k == strata - 1 else per
                res = torch.zeros((cnt, 40), dtype=torch.uint8, device=dev)
                check(sp.lib.xq_playout(b.board[lo:].data_ptr(), b.meta[lo:].data_ptr(), b.pos_hist[lo:].data_ptr(),
                                        b.hist_cap, SEED, lo, step * k, 0, res.data_ptr(), None, None, None, None,
                                        None, None, cnt, torch.cuda.current_stream().cuda_stream))
        start()
        sp.play()                                   # warm-up batch (shapes of every bucket size)
        start()
        sp.mcts.rows_evaluated = 0
        torch.cuda.synchronize()
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sp.play()
        b2.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b2)
        st = sp.stats()
        waves = (sims + 7) // 8
        out["compact" if compact else "full_batch"] = {
            "ms": ms, "batch_plies": sp.plies, "game_plies": st["plies"], "sims_per_s": st["sims"] / (ms * 1e-3),
            "rows_evaluated": int(sp.mcts.rows_evaluated), "live_rows": st["plies"] * waves,
            "rows_evaluated_per_live_row": sp.mcts.rows_evaluated / max(1, st["plies"] * waves)}
    out["speedup"] = out["full_batch"]["ms"] / out["compact"]["ms"]
    return out


def measure_mcts_multi(torch, dist, dev, world, rank, games=16384, sims=50, min_seconds=2.0, peaks=None):
    """cfg 4: 16,384 games per GPU x 50 sims/move on every rank, weights broadcast by NCCL once
    (the per-iteration collective), no collective inside the game loop, then the per-iteration
    sample gather to rank 0.  All ranks call this."""
    from chinesechessai_b200 import dist as xd
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.samples import training_tensors
    from chinesechessai_b200.self_play import BatchedSelfPlay
    torch.manual_seed(rank)                       # different weights until the broadcast
    net = ChessNet().to(dev).eval()
    xd.broadcast_weights(net, src=0)              # warm-up (NCCL communicator setup, flat buffer)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    reps = 5
    a.record()
    for _ in range(reps):
        sent = xd.broadcast_weights(net, src=0)
    b.record()
    torch.cuda.synchronize()
    bcast_ms = a.elapsed_time(b) / reps
    sp = BatchedSelfPlay(make_evaluator(torch, net, "bf16"), games, sims, temperature=1.0, device=dev,
                         seed=0, first_game_id=rank * games)
    sp.restart(SEED, MCTS_OPENING_PLIES)
    sp.play(2, check_done=False)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(dev.index or 0).start() if rank == 0 else None
    p0 = sp.plies
    n_plies = max(4, int(min_seconds / 0.030))    # ~30 ms per ply at this size -> >= 2 s
    n_plies = min(n_plies, PLIES - p0)
    a.record()
    sp.play(n_plies, check_done=False)
    b.record()
    torch.cuda.synchronize()
    dist.barrier()
    clk = clocks.stop() if clocks else None
    ms_play = a.elapsed_time(b)
    played = sp.rec_played[p0:sp.plies].sum().to(torch.int64)
    # per-iteration sample gather (trainer rank = 0) of what has been played so far
    mine = training_tensors(sp)
    mine["game"] = mine["game"] + rank * games
    xd.gather_samples({k: v[:16] for k, v in mine.items()}, dst=0)      # warm-up
    torch.cuda.synchronize()
    dist.barrier()
    a.record()
    got = xd.gather_samples(mine, dst=0)
    b.record()
    torch.cuda.synchronize()
    gather_ms = a.elapsed_time(b)
    row_bytes = sum(v[0:1].numel() * v.element_size() for v in mine.values())
    gathered = torch.tensor([0 if got is None else int(got["reward"].shape[0])], dtype=torch.int64, device=dev)
    t = torch.tensor([ms_play, bcast_ms, gather_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(played, op=dist.ReduceOp.SUM)
    dist.all_reduce(gathered, op=dist.ReduceOp.SUM)
    ms, played = float(t[0]), int(played)
    waves = (sims + 7) // 8
    peaks = peaks or read_peaks()
    tf = played * waves / (ms * 1e-3) * FLOP_PER_LEAF_EVAL / 1e12 / world
    return {"metric": "MCTS sims/sec", "value": played * sims / (ms * 1e-3), "unit": "sims/s",
            "n_gpus": world, "scaling": "weak", "precision": "bf16",
            "config": {"workload": f"cfg4: {games} games per GPU x {world} GPUs, {sims} sims/move, random-init "
                                   f"ChessNet broadcast from rank 0 by NCCL, T=1.0, {MCTS_OPENING_PLIES} random "
                                   "opening plies", "plies_timed": n_plies, "seconds_timed": ms * 1e-3},
            "ms_per_ply": ms / n_plies,
            "unique_leaf_evals_per_s": played * waves / (ms * 1e-3),
            "weight_broadcast": {"ms": float(t[1]), "bytes": int(sent),
                                 "GBps": int(sent) / (float(t[1]) * 1e-3) / 1e9 if float(t[1]) > 0 else None,
                                 "how": "one persistent flat buffer per dtype, in place (dist.FlatParams)"},
            "sample_gather": {"ms": float(t[2]), "rows": int(gathered), "bytes": int(gathered) * row_bytes,
                              "how": "row counts all-gathered, one packed send per rank to rank 0 only"},
            "clocks": clk,
            "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": tf, "peak": peaks["bf16_sustained"],
                         "frac": tf / peaks["bf16_sustained"], "traffic": None,
                         "note": "per GPU; sustained bf16 peak (leg timed >= 2 s)"}}


def measure_cfg5_multi(torch, dist, dev, world, rank, games_per_gpu=MCTS_GAMES, sims=MCTS_SIMS):
    """cfg 5 at N GPUs: ONE full iteration — NCCL weight broadcast, games sharded over the ranks
    (no collective inside self-play), sample gather to rank 0, the value-loss update of
    Trainer.train_network on rank 0 from device-resident samples (chinesechessai_b200.iteration).
    Wall seconds, max over ranks.  The single-GPU line times the same iteration through the
    reference's UNCHANGED trainer.py instead (`cfg5`)."""
    from chinesechessai_b200.iteration import self_play_iteration
    from chinesechessai_b200.neural_network import ChessNet
    torch.manual_seed(0)
    net = ChessNet().to(dev).eval()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    self_play_iteration(net, opt, 64 * world, sims, seed=1, net_dtype=torch.bfloat16)      # warm-up
    dist.barrier()
    it = self_play_iteration(net, opt, games_per_gpu * world, sims, seed=2, net_dtype=torch.bfloat16)
    t = torch.tensor([it["seconds"], it["self_play_s"], it["train_s"]], dtype=torch.float64, device=dev)
    c = torch.tensor([it["plies"], it["samples"]], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return {"workload": f"cfg5: one iteration, {games_per_gpu} games per GPU x {world} GPUs, {sims} sims/move, "
                        "bf16 inference copy for the search, fp32 update (min(50, samples//64) batches of 64) "
                        "on rank 0",
            "seconds": float(t[0]), "self_play_s": float(t[1]), "train_s": float(t[2]),
            "games": games_per_gpu * world, "plies": int(c[0]), "samples_on_trainer": int(c[1]),
            "games_per_s": games_per_gpu * world / float(t[0]), "loss": it["loss"]}


def measure_selfplay_iteration(torch, dev, games=MCTS_GAMES, sims=MCTS_SIMS):
    """The self-play half of cfg 5: `games` complete games (to a terminal state or the 70-ply cap,
    all rules active) with `sims` simulations per move through the drop-in batch loop, plus the
    training tensors Trainer.train_network consumes (boards, shaped rewards) left on the device.
    The batch is diversified by 4 random opening plies (with 15 sims the visit distribution of a
    fresh search is a delta, so identical starts would give 4,096 copies of one game).  Timed by
    wall clock around the whole call, host control flow included."""
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.samples import training_tensors
    from chinesechessai_b200.self_play import BatchedSelfPlay
    torch.manual_seed(0)
    net = ChessNet().to(dev).eval()
    ev = make_evaluator(torch, net, "bf16")
    warm = BatchedSelfPlay(ev, games, sims, temperature=1.0, device=dev, seed=1)
    warm.restart(SEED + 1, MCTS_OPENING_PLIES)
    warm.play()
    training_tensors(warm)
    del warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sp = BatchedSelfPlay(ev, games, sims, temperature=1.0, device=dev, seed=0)
    sp.restart(SEED, MCTS_OPENING_PLIES)
    sp.play()
    smp = training_tensors(sp)
    n_samples = int(smp["reward"].shape[0])
    checksum = float(smp["reward"].sum())           # device -> host read of the result
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = sp.stats()
    meta = sp.boards.meta_host()
    import numpy as np
    distinct = len(np.unique(sp.boards.position_hash().cpu().numpy()))
    return {"workload": f"self-play half of cfg 5: {games} complete games, {sims} sims/move, random-init "
                        f"ChessNet (bf16), T=1.0, {MCTS_OPENING_PLIES} random opening plies per game, samples "
                        "left on the device as training tensors",
            "seconds": dt, "games_per_s": games / dt, "plies": st["plies"], "sims_per_s": st["sims"] / dt,
            "samples": n_samples, "reward_checksum": checksum,
            "decisive_games": int(((meta["winner"] == 1) | (meta["winner"] == -1)).sum()),
            "distinct_final_positions": int(distinct),
            "reference_note": "the unmodified reference plays one such game in ~32 s per core "
                              "(BASELINE.md section 2; cfg5 below times it on this host)"}


def _run_driver(env, argv, timeout):
    drv = os.path.join(ROOT, "tests", "drivers", "drive_consumers.py")
    with tempfile.TemporaryDirectory(prefix="xq_cfg5_") as tmp:
        p = subprocess.run([sys.executable, drv] + argv, env=env, cwd=tmp, capture_output=True, text=True,
                           timeout=timeout)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    if p.returncode != 0 or not lines:
        return {"error": (p.stderr or p.stdout)[-600:]}
    return json.loads(lines[-1])


def measure_cfg5(games_ours=(100, 4096), sims=15, ref_games=4, skip_reference=False):
    """cfg 5 — one full iteration through the reference's UNCHANGED trainer.py:
    parallel_self_play(num_workers=4) -> ReplayBuffer.push -> Trainer.train_network
    (self_play.py:368, trainer.py:27-33, :298-362).  Ours: the three shim modules of
    INTEGRATION.md put the CUDA engine under that code (integration/shims); reference: the stock
    checkout with CUDA hidden (config.DEVICE == "cpu"), 4 worker processes, a sub-sampled game
    count.  Wall-clock seconds of the whole iteration, both timed on this host in this run."""
    from baseline import reference as R
    if R.locate() is None:
        return {"unavailable": "no reference checkout on this box (baseline/_ref, XQ_REFERENCE)"}
    out = {"definition": "parallel_self_play(network, G, temperature=1.0, num_simulations=S, num_workers=4) + "
                         "replay push + Trainer.train_network() [min(50, len(buffer)//64) batches of 64]",
           "sims": sims, "ours": []}
    for g in games_ours:
        r = _run_driver(R.env_for_shims(), ["--mode", "cfg5", "--games", str(g), "--sims", str(sims),
                                            "--workers", "4", "--warm", str(min(g, 64))], 900)
        if "error" not in r:
            r = {k: r[k] for k in ("games", "plies", "seconds", "self_play_s", "push_s", "train_s",
                                   "train_batches", "decisive", "engine", "device")} | \
                {"games_per_s": r["games"] / r["seconds"], "sims_per_s": r["plies"] * sims / r["self_play_s"]}
        out["ours"].append(r)
    if not skip_reference:
        cores = os.cpu_count() or 4
        r = _run_driver(R.env_for_reference(hide_cuda=True, threads=max(1, cores // 4)),
                        ["--mode", "cfg5", "--games", str(ref_games), "--sims", str(sims), "--workers", "4"], 1500)
        if "error" not in r:
            r = {k: r[k] for k in ("games", "plies", "seconds", "self_play_s", "train_s", "train_batches",
                                   "engine", "device", "cores", "torch_threads")} | \
                {"games_per_s": r["games"] / r["seconds"], "sims_per_s": r["plies"] * sims / r["self_play_s"],
                 "kind": "reference", "workers": 4,
                 "sample": f"{ref_games} games (sub-sampled; the reference's stock iteration is 100 games), "
                           "unmodified self_play.py / trainer.py, CUDA hidden"}
        out["reference"] = r
        ok = [o for o in out["ours"] if "error" not in o]
        if ok and "error" not in r:
            out["games_per_s_ratio"] = {str(o["games"]): o["games_per_s"] / r["games_per_s"] for o in ok}
    return out


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from chinesechessai_b200 import _lib
    from chinesechessai_b200.engine import BoardBatch, playout_host, results_host
    from chinesechessai_b200._lib import BOARD_STRIDE, META_DTYPE, PLAYOUT_RESULT_DTYPE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference "
                         "for the reference's CPU implementation")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    n = args.boards
    first_id = rank * n
    peaks = read_peaks()

    bb = BoardBatch(n, device=dev, hist_cap=PLIES + 2)
    results = torch.zeros((n, 40), dtype=torch.uint8, device=dev)
    res0 = torch.zeros_like(results)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def one_step(k):
        bb.reset()
        bb.playout(SEED + k, PLIES, first_game_id=first_id, results=results)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(max(args.warmup, 0)):
        one_step(1000 + w)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = lib.xq_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    plies = torch.zeros((), dtype=torch.int64, device=dev)
    barrier()
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (untimed)
        ev[k][0].record()
        bb.reset()
        ev[k][1].record()                   # kernel-only window starts after the reset launch
        bb.playout(SEED + k, PLIES, first_game_id=first_id, results=results)
        ev[k][2].record()
        plies += results.view(torch.int32)[:, 0].sum()
        if k == 0:
            res0.copy_(results)             # kept for the self-check below (outside the step's events)
    barrier()
    launches = lib.xq_launch_count() - launches0
    step_ms = sum(a.elapsed_time(c) for a, _, c in ev)
    kern_ms = sum(b.elapsed_time(c) for _, b, c in ev)
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=dev)
    total_plies = plies.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_plies, op=dist.ReduceOp.SUM)
    step_ms, kern_ms = float(t[0]), float(t[1])
    total_plies = int(total_plies)
    value = total_plies / (step_ms * 1e-3)
    clk = clocks.stop() if rank == 0 else None

    # ---- e2e: host buffers through xq_playout_host (H2D + kernel + D2H timed) ----------
    init = BoardBatch(n, device=dev, hist_cap=1)
    board0 = init.board.cpu().numpy()
    meta0 = init.meta_host()
    hb = torch.empty((n, BOARD_STRIDE), dtype=torch.int8).pin_memory()
    hm = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
    hr = torch.empty((n, 40), dtype=torch.uint8).pin_memory()
    hbn, hmn = hb.numpy(), hm.numpy().view(META_DTYPE).reshape(n)
    hrn = hr.numpy().view(PLAYOUT_RESULT_DTYPE).reshape(n)
    e2e_steps = max(1, args.steps)
    for w in range(2):
        hbn[:] = board0
        hmn[:] = meta0
        playout_host(hbn, hmn, SEED + 2000 + w, PLIES, first_id, 0, local, hrn)
    barrier()
    e2e_t, e2e_plies, e2e_res0 = 0.0, 0, None
    for k in range(e2e_steps):
        hbn[:] = board0
        hmn[:] = meta0
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        playout_host(hbn, hmn, SEED + k, PLIES, first_id, 0, local, hrn)  # synchronous call
        e2e_t += time.perf_counter() - t0
        e2e_plies += int(hrn["plies"].sum())
        if k == 0:
            e2e_res0 = hrn.copy()
    e2e_launches = e2e_steps
    et = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    ep = torch.tensor([e2e_plies], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        dist.all_reduce(ep, op=dist.ReduceOp.SUM)
    e2e_value = int(ep) / float(et)
    # host-buffer path and device-resident path of step 0 must agree on every rank
    dev_res0 = results_host(res0)
    if same_results(dev_res0, e2e_res0):
        raise SystemExit(f"bench.py rank {rank}: xq_playout_host and xq_playout disagree on step 0")

    mc_multi = cfg5_multi = None
    if world > 1 and not args.fast:
        mc_multi = measure_mcts_multi(torch, dist, dev, world, rank, peaks=peaks)
        cfg5_multi = measure_cfg5_multi(torch, dist, dev, world, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baselines + self-check of what was timed -----------------------------------------
    verified, cpu_baseline = {}, None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, _, _, _ = cpu_port(1024, threads)
        games = n if rate * 30.0 / 69.0 >= n else int(max(1024, rate * 12.0 / 69.0))
        rate, plies_c, dt, ores = cpu_port(games, threads, SEED, first_id)       # the games of timed step 0
        bad = same_results(dev_res0[:games], ores)
        verified["oracle_port"] = {"equal": not bad, "games": games, "of": n, "step": 0,
                                   "fields": list(RESULT_FIELDS) + ["reward_sum"], "fields_differing": bad}
        if bad:
            raise SystemExit(f"bench.py: timed step 0 differs from the oracle on {bad}")
        port = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{games} of {n} games of timed step 0, {plies_c} plies in {dt:.1f} s "
                          "(oracle/xq_oracle.c, pthreads)"}
        ref = None if (args.no_python or world > 1) else py_reference(12.0, SEED, first_id)
        if ref is not None:
            k = len(ref["results"])
            bad = same_results(dev_res0[:k], ref["results"])
            verified["python_reference"] = {"equal": not bad, "games": k, "step": 0, "fields_differing": bad}
            if bad:
                raise SystemExit(f"bench.py: timed step 0 differs from the Python reference on {bad}")
            cpu_baseline = {kk: ref[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
            cpu_baseline["port"] = port
        else:
            cpu_baseline = dict(port)
            cpu_baseline["reference_unavailable"] = "Python reference not timed (absent, --no-python or N>1)"

    kc = kernel_counts()
    mode = os.environ.get("XQ_PLAYOUT_MODE") or ("pairs" if n >= 24576 else "warp")
    kinfo = kc.get(mode, {})
    plies_per_launch = total_plies / (args.steps * world)
    kern_s = kern_ms * 1e-3 / args.steps
    kern_steps_per_s = plies_per_launch / kern_s
    sm_hz = (clk.get("sm_mhz") or 1965.0) * 1e6 if clk else 1965.0e6
    issue_peak = 148 * 4 * sm_hz
    wips = kinfo.get("warp_inst_per_board_step")
    hbm_achieved = BYTES_PER_STEP_FUSED * plies_per_launch / kern_s / 1e9
    roofline = {
        "bound": "issue", "unit": "warp-inst/s", "peak": issue_peak,
        "achieved": kern_steps_per_s * wips if wips else None,
        "frac": kern_steps_per_s * wips / issue_peak if wips else None,
        "warp_inst_per_board_step": wips,
        "traffic": kinfo.get("dram_bytes_per_launch") if n == BOARDS else None,
        "kernel": kinfo.get("kernel", mode), "kernel_ms_per_launch": kern_ms / args.steps,
        "source": kinfo.get("source", "no ncu capture registered for this kernel in profiles/kernel_counts.json"),
        "peak_def": "148 SMs x 4 schedulers x SM clock sampled during the timed region (one warp-instruction "
                    "per scheduler per cycle)",
        "why": "integer, branchy, tiny-state kernel: HBM traffic is ~1e-4 of the peak, SM issue slots bind "
               "(SURVEY §8d)",
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": hbm_achieved / peaks["hbm_gbs"], "bytes_per_board_step": BYTES_PER_STEP_FUSED,
                "algorithmic_bytes_per_launch": BYTES_PER_STEP_FUSED * plies_per_launch,
                "peak_source": peaks["source"]}}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8/f64", "data": "synthetic",
        "config": workload_config(n, world),
        "timing": {"l2": "flushed between timed iterations (256 MiB write, untimed)",
                   "clock": "CUDA events on the launching stream, max over ranks"},
        "plies_per_step": total_plies / args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * (BOARD_STRIDE + 32),
                "d2h_bytes_per_step": n * (BOARD_STRIDE + 32 + 40), "steps": e2e_steps,
                "api": "xq_playout_host (pinned host buffers)"},
        "gpu_launches": int(launches + e2e_launches),
        "roofline": roofline, "clocks": clk, "verified": verified,
    }
    if cpu_baseline is not None:
        out["cpu_baseline"] = cpu_baseline
    if mc_multi is not None:
        out["mcts_cfg4"] = mc_multi
    if cfg5_multi is not None:
        out["cfg5"] = cfg5_multi
    if world == 1 and not args.fast:
        # cfg 1 (the reference's own CPU-runnable case): 1,024 games from the initial position
        b1 = BoardBatch(1024, device=dev, hist_cap=PLIES + 2)
        r1 = torch.zeros((1024, 40), dtype=torch.uint8, device=dev)
        b1.playout(SEED, PLIES, results=r1)
        torch.cuda.synchronize()
        c_ms, c_plies = 0.0, 0
        for k in range(5):
            b1.reset()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            b1.playout(SEED + k, PLIES, results=r1)
            b.record()
            torch.cuda.synchronize()
            c_ms += a.elapsed_time(b)
            c_plies += int(r1.view(torch.int32)[:, 0].sum())
        out["cfg1"] = {"workload": "cfg1: 1,024 games x <= 70 plies (latency-bound: 128 CTAs on 148 SMs)",
                       "value": c_plies / (c_ms * 1e-3), "unit": UNIT, "ms_per_batch": c_ms / 5}
        if cpu_baseline is not None and cpu_baseline.get("kind") == "reference":
            # cfg 1 next to the Python reference of the SAME run: its cost per game does not depend
            # on how many games are played, so its rate on this run's sample is its cfg 1 rate
            out["cfg1"]["cpu_reference_same_run"] = {
                "value": cpu_baseline.get("value"), "unit": UNIT, "cores": cpu_baseline.get("cores"),
                "kind": "reference", "sample": cpu_baseline.get("sample")}
        out["gpu_launches"] += 11
        # the other mappings of the same fused loop
        out["other_mappings"] = {}
        for m2 in ("pairs", "pair", "pairq", "tpb", "warp"):
            if m2 == mode:
                continue
            os.environ["XQ_PLAYOUT_MODE"] = m2
            one_step(3000)
            torch.cuda.synchronize()
            w_ms, w_plies = 0.0, 0
            for k in range(3):
                flush.zero_()
                bb.reset()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                bb.playout(SEED + k, PLIES, first_game_id=first_id, results=results)
                b.record()
                torch.cuda.synchronize()
                w_ms += a.elapsed_time(b)
                w_plies += int(results.view(torch.int32)[:, 0].sum())
            del os.environ["XQ_PLAYOUT_MODE"]
            wv = w_plies / (w_ms * 1e-3)
            wi = kc.get(m2, {}).get("warp_inst_per_board_step")
            out["other_mappings"][m2] = {"value": wv, "unit": UNIT, "kernel_ms_per_launch": w_ms / 3,
                                         "kernel": kc.get(m2, {}).get("kernel", m2),
                                         "issue_frac": wv * wi / issue_peak if wi else None}
            out["gpu_launches"] += 8
        l_sp = lib.xq_launch_count()
        out["selfplay_iteration"] = measure_selfplay_iteration(torch, dev)
        out["selfplay_iteration"]["gpu_launches"] = int(lib.xq_launch_count() - l_sp)
        v, ms, launches_per_step, final = measure_step_per_launch(torch, BoardBatch, n, first_id, 2, flush)
        bb.reset()
        bb.playout(SEED + 1, PLIES, first_game_id=first_id, results=results)
        same = bool(np.array_equal(final[0], bb.meta_host()["move_count"]) and
                    np.array_equal(final[1], bb.boards_host()))
        if not same:
            raise SystemExit("bench.py: step-per-launch and fused playout end in different states")
        out["step_per_launch"] = {
            "value": v, "unit": UNIT, "ms_per_step": ms, "launches_per_step": launches_per_step,
            "how": "xq_step_pick: ONE launch per ply, 70 launches replayed from a CUDA graph",
            "equals_fused_playout": same,
            "roofline": {"bound": "hbm", "achieved": v * BYTES_PER_STEP_LAUNCH_MODE / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": v * BYTES_PER_STEP_LAUNCH_MODE / 1e9 / peaks["hbm_gbs"],
                         "bytes_per_board_step": BYTES_PER_STEP_LAUNCH_MODE, "traffic": None}}
        out["gpu_launches"] += 3 * launches_per_step
        # ---- MCTS legs: every precision, each timed >= 2 s with its own clock samples ---------
        l0 = lib.xq_launch_count()
        from chinesechessai_b200.neural_network import ChessNet
        torch.manual_seed(0)
        out["mcts_precision_agreement"] = precision_agreement(torch, dev, ChessNet().to(dev).eval())
        out["mcts"] = {p: measure_mcts(torch, dev, p, peaks=peaks) for p in PRECISIONS}
        out["mcts"]["tree_only"] = measure_tree_only(torch, dev, MCTS_GAMES, MCTS_SIMS)
        out["mcts"]["ragged_batch"] = measure_mcts_ragged(torch, dev)
        if not args.no_cfg4:
            out["mcts_cfg4"] = {p: measure_mcts(torch, dev, p, games=16384, sims=50,
                                                label="cfg4 (one GPU's shard)", peaks=peaks)
                                for p in PRECISIONS}
        out["mcts_gpu_launches"] = int(lib.xq_launch_count() - l0)
        if not args.no_cfg5:
            out["cfg5"] = measure_cfg5(skip_reference=args.no_python)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--boards", type=int, default=BOARDS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs and the self-check")
    ap.add_argument("--no-python", action="store_true", help="skip the legs that run the unmodified Python reference")
    ap.add_argument("--fast", action="store_true", help="skip the step-per-launch, MCTS and cfg5 legs")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the 16,384-game x 50-sim MCTS legs")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the full-iteration leg (unchanged trainer.py)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its
    # version banner on the first communicator), so everything but that line goes to stderr:
    # fd 1 points at stderr while the run is in progress and is restored for the final print.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
    for ln in lines[:-1]:
        print(ln, file=sys.stderr)
    if lines:
        print(lines[-1], flush=True)


if __name__ == "__main__":
    main()

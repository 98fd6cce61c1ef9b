#!/usr/bin/env python
"""bench.py — board-steps/s of the Xiangqi rules hot path on B200 (BASELINE.json metric).

Workload (N=1): BASELINE.json configs[1] — batched random playouts, 65,536 boards x <=70 plies
from the initial position, uniform legal move per ply by the shared counter-based pick rule
(philox4x32-10 keyed by seed, counter (game_id, ply)); one "step" = one such batch through
the fused playout kernel (get_legal_moves + make_move per ply, every terminal rule).  N>1:
the same batch per GPU (weak scaling), game ids offset per rank, no data-path collective.

    python bench.py --gpus N --steps K --warmup W          # our arm
    python bench.py --impl reference ...                   # CPU port of the reference path

Prints ONE JSON line (see the keys at the bottom).  ``value`` is device-resident throughput,
``e2e`` goes through the host-buffer C-ABI call (H2D + kernel + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BOARDS = 65536
PLIES = 70
SEED = 0x5EED
# Algorithmic HBM bytes per board-step (DESIGN.md §Measurement / SURVEY.md §8d):
#   step-per-launch mode: state in 128 + state out 128 + legal list 81 + move 4 + history 8
BYTES_PER_STEP_LAUNCH_MODE = 352
#   fused playout: (board+meta in 128, out 128, result 40) / 70 plies + one 8-byte history append
BYTES_PER_STEP_FUSED = (128 + 128 + 40) / 70.0 + 8.0
# warp-instructions per board-step from ncu smsp__inst_executed.sum / plies of the same launch,
# refreshed with every capture: the shipped pair-per-board kernel
# (profiles/r1/playout_pair_ncu_summary.txt), the thread-per-board kernel
# (profiles/r1/playout_tpb_ncu_summary.txt) and the warp-per-board kernel
# (profiles/r1/playout_v7_ncu_summary.txt)
WARP_INST_PER_STEP = 1092.0
WARP_INST_PER_STEP_TPB_MODE = 956.0
WARP_INST_PER_STEP_WARP_MODE = 2193.0
# dram__bytes_read.sum + dram__bytes_write.sum of one playout launch in the same captures
DRAM_TRAFFIC_PER_LAUNCH = 45.86e6 + 5.39e6
FLOP_PER_LEAF_EVAL = 263_209_216          # ChessNet.forward, SURVEY.md §8d
MCTS_GAMES, MCTS_SIMS, MCTS_OPENING_PLIES = 4096, 15, 4
METRIC = "board-steps/sec (legal movegen+step)"
UNIT = "board-steps/s"


def read_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def cpu_port_rate(n_games: int, threads: int, seed: int = SEED, first: int = 0):
    """Oracle (C port of the reference path) on the host cores: plies/s over n_games playouts."""
    from oracle import xq_oracle as xo
    xo.build()
    t0 = time.perf_counter()
    total, _ = xo.playout_many(n_games, seed, first, PLIES, 0, n_threads=threads)
    dt = time.perf_counter() - t0
    return total / dt, total, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the path (the reference is pure
    Python and cannot travel to the GPU box, so this is the C port in oracle/, kind "port"),
    all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rate, _, _ = cpu_port_rate(1024, threads)                   # calibrate
    games = int(max(256, min(args.boards, rate * 8.0 / 69.0)))  # ~8 s per step
    for w in range(args.warmup):
        cpu_port_rate(min(games, 512), threads, SEED + 1000 + w)
    tot_plies, tot_t = 0, 0.0
    for k in range(args.steps):
        r, plies, dt = cpu_port_rate(games, threads, SEED + k)
        tot_plies += plies
        tot_t += dt
    value = tot_plies / tot_t
    sample = f"{games} of {args.boards} games per step from the initial position, <= {PLIES} plies"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8/f64",
        "data": "synthetic", "config": {"workload": f"cfg2 random playouts {args.boards}x{PLIES} (sampled)",
                                        "pick": "philox4x32-10(seed,(game,ply)) mod n_legal"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def measure_step_per_launch(torch, BoardBatch, n, first_id, steps, flush):
    """API-faithful mode: one xq_pick_moves + one xq_step (which also returns the next legal
    list) per ply, state round-trips through HBM every ply (352 algorithmic bytes/board-step)."""
    bb = BoardBatch(n, hist_cap=PLIES + 2)
    mv = torch.empty((n,), dtype=torch.int16, device=bb.device)

    def one(k):
        bb.reset()
        bb.legal_moves()
        for ply in range(PLIES):
            bb.pick(SEED + k, ply, first_game_id=first_id, out=mv)
            bb.step(mv, want_next=True)
    one(1000)
    torch.cuda.synchronize()
    ms, plies = 0.0, 0
    for k in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one(k)
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
        plies += int(bb.meta_host()["move_count"].sum())
    return plies / (ms * 1e-3), ms / steps, 2 * PLIES + 2


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
        while time.perf_counter() - t0 < 4.0:       # bounded sample: ~4 s of host work
            list(ex.map(one, range(n_cpu)))
            done += n_cpu
    dt = time.perf_counter() - t0
    n_cpu = done
    return {"gpu_sims_per_s": gpu,
            "cpu_baseline": {"value": n_cpu * sims / dt, "unit": "sims/s", "cores": threads, "kind": "port",
                             "sample": f"{n_cpu} searches x {sims} sims, literal replay per simulation "
                                       f"(oracle/xq_oracle.c via ctypes threads), {dt:.1f} s"}}


def measure_mcts_multi(torch, dist, dev, world, rank, games=16384, sims=50, plies_timed=2):
    """cfg 4: 16,384 games per GPU x 50 sims/move on every rank, weights broadcast by NCCL once
    (the per-iteration collective), no collective inside the game loop.  All ranks call this."""
    from chinesechessai_b200 import dist as xd
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.self_play import BatchedSelfPlay
    torch.manual_seed(rank)                       # different weights until the broadcast
    net = ChessNet().to(dev).eval()
    xd.broadcast_weights(net, src=0)              # warm-up (NCCL communicator setup)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    sent = xd.broadcast_weights(net, src=0)
    b.record()
    torch.cuda.synchronize()
    bcast_ms = a.elapsed_time(b)
    sp = BatchedSelfPlay(net, games, sims, temperature=1.0, device=dev, net_dtype=torch.bfloat16,
                         seed=0, first_game_id=rank * games)
    sp.boards.playout(SEED, MCTS_OPENING_PLIES, first_game_id=rank * games)
    sp.play(2, check_done=False)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    p0 = sp.plies
    a.record()
    sp.play(plies_timed, check_done=False)
    b.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([a.elapsed_time(b), bcast_ms], dtype=torch.float64, device=dev)
    played = sp.rec_played[p0:sp.plies].sum().to(torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(played, op=dist.ReduceOp.SUM)
    ms, played = float(t[0]), int(played)
    waves = (sims + 7) // 8
    return {"metric": "MCTS sims/sec", "value": played * sims / (ms * 1e-3), "unit": "sims/s",
            "n_gpus": world, "scaling": "weak",
            "config": {"workload": f"cfg4: {games} games per GPU x {world} GPUs, {sims} sims/move, random-init "
                                   f"ChessNet broadcast from rank 0 by NCCL, T=1.0, {MCTS_OPENING_PLIES} random "
                                   "opening plies", "plies_timed": plies_timed,
                       "nn_dtype": "bf16 (BN folded, fused cuDNN conv+bias+ReLU; reference: fp32)"},
            "ms_per_ply": ms / plies_timed,
            "unique_leaf_evals_per_s": played * waves / (ms * 1e-3),
            "weight_broadcast": {"ms": float(t[1]), "bytes": int(sent)},
            "roofline": {"bound": "tensor", "unit": "TFLOP/s",
                         "achieved": played * waves / (ms * 1e-3) * FLOP_PER_LEAF_EVAL / 1e12 / world,
                         "note": "per GPU"}}


def measure_mcts(torch, dev, plies_timed=6, games=MCTS_GAMES, sims=MCTS_SIMS, label="cfg3"):
    """cfg 3: 4,096 concurrent self-play games, 15 sims/move (2 waves of 8+7), random-init ChessNet
    (torch.manual_seed(0)), temperature 1.0; games diversified by 4 random opening plies.
    cfg 4 (per GPU): 16,384 games, 50 sims/move (7 waves)."""
    MCTS_GAMES, MCTS_SIMS = games, sims
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.self_play import BatchedSelfPlay
    torch.manual_seed(0)
    net = ChessNet().to(dev).eval()
    sp = BatchedSelfPlay(net, MCTS_GAMES, MCTS_SIMS, temperature=1.0, device=dev,
                         net_dtype=torch.bfloat16, seed=0)
    sp.boards.playout(SEED, MCTS_OPENING_PLIES)          # diversify the batch
    sp.play(3, check_done=False)                         # warm-up plies (cuDNN autotune etc.)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0 = sp.plies
    a.record()
    sp.play(plies_timed, check_done=False)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    graph_launches = plies_timed * sp.graph_kernels if sp.use_graph else 0
    played = int(sp.rec_played[p0:sp.plies].sum())
    sims = played * MCTS_SIMS
    waves = (MCTS_SIMS + 7) // 8
    evals = played * waves
    return {"metric": "MCTS sims/sec", "value": sims / (ms * 1e-3), "unit": "sims/s",
            "config": {"workload": f"{label}: {MCTS_GAMES} concurrent games, {MCTS_SIMS} sims/move, "
                                   f"random-init ChessNet, T=1.0, {MCTS_OPENING_PLIES} random opening plies",
                       "plies_timed": plies_timed, "nn_dtype": "bf16 autocast (reference: fp32)",
                       "cuda_graph": bool(sp.use_graph)},
            "xq_kernels_replayed_by_graph": graph_launches,
            "ms_per_ply": ms / plies_timed, "unique_leaf_evals_per_s": evals / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": evals / (ms * 1e-3) * FLOP_PER_LEAF_EVAL / 1e12,
                         "unit": "TFLOP/s", "flop_per_leaf_eval": FLOP_PER_LEAF_EVAL}}


def measure_selfplay_iteration(torch, dev, games=MCTS_GAMES, sims=MCTS_SIMS):
    """The self-play half of cfg 5: `games` complete games (to a terminal state or the 70-ply cap,
    all rules active) with `sims` simulations per move through the drop-in batch loop, plus the
    training tensors Trainer.train_network consumes (boards, shaped rewards) left on the device.
    Timed by wall clock around the whole call, host control flow included."""
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.samples import training_tensors
    from chinesechessai_b200.self_play import BatchedSelfPlay
    torch.manual_seed(0)
    net = ChessNet().to(dev).eval()
    # one complete untimed iteration first: allocator growth, library handles and the first-call
    # costs of the sample-shaping ops are warm-up, like the W warm-up steps of the main leg
    warm = BatchedSelfPlay(net, games, sims, temperature=1.0, device=dev, net_dtype=torch.bfloat16, seed=1)
    warm.play()
    training_tensors(warm)
    del warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sp = BatchedSelfPlay(net, games, sims, temperature=1.0, device=dev, net_dtype=torch.bfloat16, seed=0)
    sp.play()
    smp = training_tensors(sp)
    n_samples = int(smp["reward"].shape[0])
    checksum = float(smp["reward"].sum())           # device -> host read of the result
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = sp.stats()
    meta = sp.boards.meta_host()
    return {"workload": f"self-play half of cfg 5: {games} complete games, {sims} sims/move, random-init "
                        "ChessNet (bf16), T=1.0, samples left on the device as training tensors",
            "seconds": dt, "games_per_s": games / dt, "plies": st["plies"], "sims_per_s": st["sims"] / dt,
            "samples": n_samples, "reward_checksum": checksum,
            "decisive_games": int((meta["winner"] != 2).sum() - (meta["winner"] == 0).sum()),
            "reference_note": "the unmodified reference plays one such game in 32 s per core "
                              "(BASELINE.md section 2)"}


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
                         "for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    n = args.boards
    first_id = rank * n

    bb = BoardBatch(n, device=dev, hist_cap=PLIES + 2)
    results = torch.zeros((n, 40), dtype=torch.uint8, device=dev)
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
    e2e_t, e2e_plies = 0.0, 0
    for k in range(e2e_steps):
        hbn[:] = board0
        hmn[:] = meta0
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        playout_host(hbn, hmn, SEED + k, PLIES, first_id, 0, local, hrn)  # synchronous call
        e2e_t += time.perf_counter() - t0
        e2e_plies += int(hrn["plies"].sum())
    e2e_launches = e2e_steps
    et = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    ep = torch.tensor([e2e_plies], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        dist.all_reduce(ep, op=dist.ReduceOp.SUM)
    e2e_value = int(ep) / float(et)
    # device result of step 0 == host-path result of step 0 (same seed): cheap self-check
    clk = clocks.stop() if rank == 0 else None
    mc_multi = None
    if world > 1 and not args.fast:
        mc_multi = measure_mcts_multi(torch, dist, dev, world, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = read_peaks()
    plies_per_launch = total_plies / (args.steps * world)
    kern_s = kern_ms * 1e-3 / args.steps
    achieved = BYTES_PER_STEP_FUSED * plies_per_launch / kern_s / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8/f64", "data": "synthetic",
        "config": {"workload": f"cfg2: batched random playouts, {n} boards x <= {PLIES} plies per GPU "
                               "from the initial position, fused playout kernel",
                   "boards_per_gpu": n, "max_plies": PLIES,
                   "pick": "philox4x32-10(seed,(game,ply)) mod n_legal",
                   "l2": "flushed between timed iterations (256 MiB write)",
                   "parallelism": f"games sharded over {world} GPU(s), no collective on the path"},
        "plies_per_step": total_plies / args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * (BOARD_STRIDE + 32),
                "d2h_bytes_per_step": n * (BOARD_STRIDE + 32 + 40), "steps": e2e_steps,
                "api": "xq_playout_host (pinned host buffers)"},
        "gpu_launches": int(launches + e2e_launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak,
                     "traffic": DRAM_TRAFFIC_PER_LAUNCH if n == BOARDS else None,
                     "traffic_note": "ncu dram bytes read+write per launch (profiles/r1/playout_pair_ncu_summary.txt); "
                                     "algorithmic bytes per launch = %.2e" % (BYTES_PER_STEP_FUSED * plies_per_launch),
                     "peak_source": peak_src,
                     "kernel": "xq::playout_lane_kernel<false,true> (two lanes per board)" if n >= 40960
                               else "xq::playout_kernel<false,4,32> (one warp per board)",
                     "bytes_per_board_step": BYTES_PER_STEP_FUSED,
                     "kernel_ms_per_launch": kern_ms / args.steps,
                     "note": "integer/latency-bound kernel: SM issue rate binds, not HBM "
                             "(see profiles/ and DESIGN.md)"},
        "clocks": clk,
    }
    sm_hz = (clk.get("sm_mhz") or 1965.0) * 1e6 if clk else 1965.0e6
    issue_peak = 148 * 4 * sm_hz
    kern_steps_per_s = plies_per_launch / kern_s
    wips_main = WARP_INST_PER_STEP if n >= 40960 else WARP_INST_PER_STEP_WARP_MODE
    out["issue"] = {"achieved": kern_steps_per_s * wips_main, "peak": issue_peak,
                    "unit": "warp-inst/s", "frac": kern_steps_per_s * wips_main / issue_peak,
                    "warp_inst_per_board_step": wips_main,
                    "source": "ncu smsp__inst_executed.sum / plies (profiles/), peak = 148 SM x 4 x f_SM"}
    if mc_multi is not None:
        try:
            tfp = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        except Exception:
            tfp = 1400.0
        mc_multi["roofline"]["peak"] = tfp
        mc_multi["roofline"]["frac"] = mc_multi["roofline"]["achieved"] / tfp
        out["mcts_cfg4"] = mc_multi
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
        out["cfg1"] = {"workload": "cfg1: 1,024 games x <= 70 plies (latency-bound: 8 CTAs on 148 SMs)",
                       "value": c_plies / (c_ms * 1e-3), "unit": UNIT, "ms_per_batch": c_ms / 5}
        out["gpu_launches"] += 11
        # the other two mappings of the same fused kernel: a warp per board (what the API kernels
        # and MCTS use; the dispatch picks it below 40,960 boards) and a thread per board
        out["other_mappings"] = {}
        for mode, kname, wips in (("warp", "xq::playout_kernel<false,4,32>", WARP_INST_PER_STEP_WARP_MODE),
                                  ("tpb", "xq::playout_lane_kernel<false,false>", WARP_INST_PER_STEP_TPB_MODE)):
            os.environ["XQ_PLAYOUT_MODE"] = mode
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
            out["other_mappings"][mode] = {"value": wv, "unit": UNIT, "kernel_ms_per_launch": w_ms / 3,
                                           "kernel": kname,
                                           "issue": {"warp_inst_per_board_step": wips,
                                                     "frac": wv * wips / issue_peak}}
            out["gpu_launches"] += 8
        l_sp = lib.xq_launch_count()
        out["selfplay_iteration"] = measure_selfplay_iteration(torch, dev)
        out["selfplay_iteration"]["gpu_launches"] = int(lib.xq_launch_count() - l_sp)
        v, ms, launches_per_step = measure_step_per_launch(torch, BoardBatch, n, first_id, 2, flush)
        out["step_per_launch"] = {
            "value": v, "unit": UNIT, "ms_per_step": ms, "launches_per_step": launches_per_step,
            "roofline": {"bound": "hbm", "achieved": v * BYTES_PER_STEP_LAUNCH_MODE / 1e9,
                         "peak": hbm_peak, "unit": "GB/s",
                         "frac": v * BYTES_PER_STEP_LAUNCH_MODE / 1e9 / hbm_peak,
                         "bytes_per_board_step": BYTES_PER_STEP_LAUNCH_MODE}}
        out["gpu_launches"] += 3 * launches_per_step
        l0 = lib.xq_launch_count()
        mc = measure_mcts(torch, dev)
        tf_peak = None
        try:
            tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        except Exception:
            tf_peak = 1400.0
        mc["roofline"]["peak"] = tf_peak
        mc["roofline"]["frac"] = mc["roofline"]["achieved"] / tf_peak
        mc["tree_only"] = measure_tree_only(torch, dev, MCTS_GAMES, MCTS_SIMS)
        mc["gpu_launches"] = int(lib.xq_launch_count() - l0) + int(mc.get("xq_kernels_replayed_by_graph", 0))
        out["mcts"] = mc
        if not args.no_cfg4:
            m4 = measure_mcts(torch, dev, plies_timed=2, games=16384, sims=50, label="cfg4 (one GPU's shard)")
            m4["roofline"]["peak"] = tf_peak
            m4["roofline"]["frac"] = m4["roofline"]["achieved"] / tf_peak
            out["mcts_cfg4"] = m4
    # ---- CPU baseline: C port of the reference path on the host cores, bounded sample ------
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, _, _ = cpu_port_rate(1024, threads)
        games = int(max(256, min(BOARDS, rate * 12.0 / 69.0)))
        rate, plies_c, dt = cpu_port_rate(games, threads)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{games} of {BOARDS} games, {plies_c} plies in {dt:.1f} s "
                                         "(oracle/xq_oracle.c, pthreads)"}
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--fast", action="store_true", help="skip the step-per-launch and MCTS legs")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the 16,384-game x 50-sim MCTS leg")
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

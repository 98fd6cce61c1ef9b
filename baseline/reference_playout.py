"""Random playouts on the UNMODIFIED Python reference (chess_env.ChineseChess), cfg 1 / cfg 2
style: get_legal_moves() + make_move() per ply from the initial position, the move chosen by the
shared counter-based pick rule (philox4x32-10 keyed by the seed, counter (game id, ply)), the same
per-ply digest chain the CUDA kernels and the C oracle compute.

Bench / test infrastructure only: it imports the reference from ``baseline.reference.locate()``
and is never imported by the product package.  The pick rule and the digest are restated here in
plain Python so that this file depends on nothing but the reference itself and numpy.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time
from typing import Dict, List, Optional, Tuple

import numpy as np

M64 = (1 << 64) - 1
RESULT_DTYPE = np.dtype([("plies", "<i4"), ("winner", "<i4"), ("reason", "<i4"), ("max_legal", "<i4"),
                         ("reward_sum", "<f8"), ("digest", "<u8"), ("final_hash", "<u8")])


def mix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def position_key(board: np.ndarray, player: int) -> int:
    """The engine's 64-bit stand-in for hash(board.tobytes() + side byte) (chess_env.py:497-504)."""
    h = mix64(0x7000 + (0 if player == 1 else 1))
    for sq, p in enumerate(board.reshape(90).tolist()):
        if p:
            h ^= mix64((p + 8) * 128 + sq)
    return h


def philox4x32(c0: int, c1: int, c2: int, c3: int, k0: int, k1: int) -> Tuple[int, int, int, int]:
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, \
            ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def pack(mv) -> int:
    return (mv[0] * 9 + mv[1]) * 90 + mv[2] * 9 + mv[3]


_REASONS = (("吃掉对方将帅", 1, "end"), ("将死", 2, "start"), ("三次重复局面判和", 3, "eq"),
            ("50回合无吃子判和", 4, "eq"), ("困毙", 5, "start"), ("长将判负", 6, "start"),
            ("长捉判负", 7, "start"), ("超过", 8, "start"))


def reason_code(s: Optional[str]) -> int:
    if s is None:
        return 0
    for text, code, how in _REASONS:
        if (how == "end" and s.endswith(text)) or (how == "start" and s.startswith(text)) or s == text:
            return code
    return 15


_chess_env = None


def _reference_env():
    global _chess_env
    if _chess_env is None:
        from baseline.reference import locate
        ref = locate()
        if ref is None:
            raise FileNotFoundError("reference checkout not found")
        sys.path.insert(0, ref)
        sys.dont_write_bytecode = True
        with contextlib.redirect_stdout(io.StringIO()):  # config.py prints the device at import
            import chess_env
        if os.path.realpath(os.path.dirname(chess_env.__file__)) != os.path.realpath(ref):
            raise ImportError(f"chess_env resolved to {chess_env.__file__}, not to the reference in {ref}")
        _chess_env = chess_env
    return _chess_env


def play(args) -> Tuple[int, int, int, int, float, int, int]:
    """One game -> (plies, winner (2 = None), reason code, max_legal, reward_sum, digest, final_hash)."""
    game_id, seed, max_plies, bias = args
    env = _reference_env().ChineseChess()
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    digest, rsum, max_legal, ply = 0, 0.0, 0, 0
    while ply < max_plies:
        legal = env.get_legal_moves()
        if not legal:
            break
        packed = [pack(m) for m in legal]
        n = len(packed)
        max_legal = max(max_legal, n)
        x = philox4x32(game_id, ply, 0, 0, k0, k1)
        idx = x[0] % n
        if bias > 0 and (x[1] & 0xFF) < bias:
            flat = env.board.reshape(90)
            caps = [i for i, m in enumerate(packed) if flat[m % 90] != 0]
            if caps:
                idx = caps[x[0] % len(caps)]
        _, reward, done = env.make_move(legal[idx])
        lsum = 0
        for i, m in enumerate(packed):
            lsum = (lsum + (m + 1) * (2 * i + 1)) & 0xFFFFFFFF
        a = lsum | (n << 32) | (packed[idx] << 40) | ((ply + 1) << 54)
        wn = 2 if env.winner is None else env.winner
        c = (1 if done else 0) | ((wn + 2) << 8) | (reason_code(env.end_reason) << 16) | \
            ((1 if isinstance(reward, int) else 0) << 24)
        word = (a * 0x9E3779B97F4A7C15 + int(np.float64(reward).view(np.uint64)) * 0xC2B2AE3D27D4EB4F +
                c * 0x165667B19E3779F9 + position_key(env.board, env.current_player) * 0x27D4EB2F165667C5) & M64
        digest = mix64(digest ^ word)
        rsum += float(reward)
        ply += 1
        if done:
            break
    wn = 2 if env.winner is None else int(env.winner)
    return (ply, wn, reason_code(env.end_reason), max_legal, rsum, digest,
            position_key(env.board, env.current_player))


def _warm(_):
    _reference_env()
    return os.getpid()


def playout_many(game_ids: List[int], seed: int, max_plies: int = 70, bias: int = 0,
                 procs: Optional[int] = None) -> Dict[str, object]:
    """Play ``game_ids`` with one process per core (multiprocessing, fork).  Returns the results
    as a RESULT_DTYPE array in the order of ``game_ids`` and the wall time of the games alone
    (the pool's start-up and the reference's import are outside the timed region)."""
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    out = np.zeros(len(game_ids), RESULT_DTYPE)
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_warm, range(procs * 2), chunksize=1)
        t0 = time.perf_counter()
        rows = pool.map(play, [(int(g), int(seed), int(max_plies), int(bias)) for g in game_ids], chunksize=1)
        dt = time.perf_counter() - t0
    for i, r in enumerate(rows):
        out[i] = r
    return {"results": out, "seconds": dt, "plies": int(out["plies"].sum()), "procs": procs}


def run_subprocess(first: int, count: int, seed: int, max_plies: int = 70, bias: int = 0,
                   procs: Optional[int] = None, timeout: float = 900.0) -> Dict[str, object]:
    """playout_many() in a FRESH interpreter with CUDA hidden: a process that has already
    initialised CUDA (bench.py's GPU arm) cannot fork workers that import the reference's
    config.py (it probes torch.cuda at import)."""
    import json
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["PYTHONPATH"] = root
    env["CUDA_VISIBLE_DEVICES"] = ""
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    with tempfile.TemporaryDirectory(prefix="xq_refplay_") as tmp:
        out = os.path.join(tmp, "res.npy")
        cmd = [sys.executable, "-m", "baseline.reference_playout", "--first", str(first), "--count", str(count),
               "--seed", str(seed), "--max-plies", str(max_plies), "--bias", str(bias), "--out", out]
        if procs:
            cmd += ["--procs", str(procs)]
        p = subprocess.run(cmd, env=env, cwd=root, capture_output=True, text=True, timeout=timeout)
        lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        if p.returncode != 0 or not lines:
            raise RuntimeError("reference playout subprocess failed: " + (p.stderr or p.stdout)[-800:])
        info = json.loads(lines[-1])
        info["results"] = np.load(out)
    return info


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=8)
    ap.add_argument("--seed", type=int, default=0x5EED)
    ap.add_argument("--max-plies", type=int, default=70)
    ap.add_argument("--bias", type=int, default=0)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    r = playout_many(list(range(a.first, a.first + a.count)), a.seed, a.max_plies, a.bias, a.procs or None)
    if a.out:
        np.save(a.out, r["results"])
    print(json.dumps({"seconds": r["seconds"], "plies": r["plies"], "procs": r["procs"], "games": a.count}))

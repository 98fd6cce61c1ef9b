"""Times the UNMODIFIED Python reference (needs /root/reference; authoring container only):
cfg-1-style random playouts, get_legal_moves + make_move per ply, shared pick rule, one process
per core.  Prints board-steps/s; the number is quoted in DESIGN.md next to the GPU results."""
import contextlib, io, multiprocessing as mp, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.dont_write_bytecode = True
REF = os.environ.get("XQ_REFERENCE", "/root/reference")


def play(game_id):
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import chess_env
    from tests.golden.gen_golden import pick_index, pack, SEED
    env = chess_env.ChineseChess()
    plies = 0
    for ply in range(70):
        legal = env.get_legal_moves()
        if not legal:
            break
        idx = pick_index(env.board, [pack(m) for m in legal], SEED, game_id, ply, 0)
        _, _, done = env.make_move(legal[idx])
        plies += 1
        if done:
            break
    return plies


if __name__ == "__main__":
    procs = os.cpu_count() or 1
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * procs
    with mp.Pool(procs) as pool:
        pool.map(play, range(procs))          # warm-up / imports
        t0 = time.perf_counter()
        plies = sum(pool.map(play, range(1000, 1000 + n), chunksize=1))
        dt = time.perf_counter() - t0
    print(f"reference (Python, unmodified): {n} games, {plies} plies in {dt:.1f} s on {procs} processes "
          f"-> {plies / dt:.1f} board-steps/s ({plies / dt / procs:.1f} per core)")

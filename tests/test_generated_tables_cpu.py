"""CPU: the generated lookup tables compiled into the kernels (chinesechessai_b200/csrc/*.inc) are
what their generator scripts produce, and hold the geometry the reference's generators encode
(chess_env.py:123-251), checked here against plain restatements of those rules."""
import os
import subprocess
import sys

import pytest

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "chinesechessai_b200", "csrc")
TABLES = ("leap", "touch", "ray", "knight")


@pytest.mark.parametrize("name", TABLES)
def test_inc_file_is_fresh(name):
    out = subprocess.run([sys.executable, os.path.join(CSRC, f"gen_{name}_table.py")], capture_output=True,
                         text=True, check=True).stdout
    assert out == open(os.path.join(CSRC, f"xq_{name}_table.inc")).read(), f"re-run gen_{name}_table.py"


def _words(name):
    txt = open(os.path.join(CSRC, f"xq_{name}_table.inc")).read()
    return [int(t.rstrip("u"), 16) for line in txt.splitlines() if not line.startswith("//")
            for t in line.replace(",", " ").split()]


def test_knight_table_geometry():
    """entry [T][diagonal] = leg | knight_a << 8 | knight_b << 16: a knight on knight_a / knight_b
    reaches T exactly when the leg (adjacent to the KNIGHT, chess_env.py:182-195) is the named one."""
    w = _words("knight")
    assert len(w) == 90 * 4
    knight_moves = [(2, 1, 1, 0), (2, -1, 1, 0), (-2, 1, -1, 0), (-2, -1, -1, 0),
                    (1, 2, 0, 1), (-1, 2, 0, 1), (1, -2, 0, -1), (-1, -2, 0, -1)]   # (dr, dc, leg_dr, leg_dc) :182-187
    for T in range(90):
        tr, tc = divmod(T, 9)
        want = set()
        for r in range(10):
            for c in range(9):
                for dr, dc, lr, lc in knight_moves:
                    if (r + dr, c + dc) == (tr, tc):
                        want.add((r * 9 + c, (r + lr) * 9 + (c + lc)))
        got = set()
        for i in range(4):
            e = w[T * 4 + i]
            leg, a, b = e & 0xFF, (e >> 8) & 0xFF, (e >> 16) & 0xFF
            if leg == 0xFF:
                assert a == 0xFF and b == 0xFF
                continue
            got |= {(k, leg) for k in (a, b) if k != 0xFF}
        assert got == want, T


def test_ray_table_entries():
    """entry = empties | first << 4 | second << 8 for every line occupancy: compared with a scan."""
    w = _words("ray")
    assert len(w) == 9 * 512 * 2 + 10 * 1024 * 2
    idx = 0
    for length, bits in ((9, 512), (10, 1024)):
        for x in range(length):
            for m in range(bits):
                for back in (0, 1):
                    seen, empties = [], None
                    pos, k = x + (-1 if back else 1), 1
                    while 0 <= pos < length and len(seen) < 2:
                        if (m >> pos) & 1:
                            if not seen:
                                empties = k - 1
                            seen.append(k)
                        pos += -1 if back else 1
                        k += 1
                    if empties is None:
                        empties = x if back else length - 1 - x
                    first = seen[0] if seen else 0
                    second = seen[1] if len(seen) > 1 else 0
                    assert w[idx] == empties | first << 4 | second << 8, (length, x, m, back)
                    idx += 1


def test_touch_table_bits():
    """entry [K][s]: row bit, column bit, leg bit and knight bit of s as seen from a king on K."""
    w = _words("touch")
    assert len(w) == 90 * 90
    for K in range(0, 90, 7):
        kr, kc = divmod(K, 9)
        for s in range(90):
            r, c = divmod(s, 9)
            e = w[K * 90 + s]
            assert (e & 0x1FF) == ((1 << c) if r == kr else 0)
            assert ((e >> 9) & 0x3FF) == ((1 << r) if c == kc else 0)
            dr, dc = r - kr, c - kc
            assert bool((e >> 19) & 0xF) == (abs(dr) == 1 and abs(dc) == 1)
            assert bool((e >> 23) & 0xFF) == (sorted((abs(dr), abs(dc))) == [1, 2])
            assert bin(e >> 19).count("1") <= 1

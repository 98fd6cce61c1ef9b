"""CPU tests of host-side logic that needs no GPU: the vectorised reward shaping and sample
materialisation of the batched game loop against the straightforward per-sample formulation of
self_play.py:262-310, and the device-side reward formula against the same reference branches."""
import pickle

import numpy as np
import torch

from chinesechessai_b200 import self_play as sp
from chinesechessai_b200.samples import _final_reward


def _random_batch(rng, P=70, n=48):
    rn = rng.integers(20, 60, size=(P, n)).astype(np.int16)
    rm = rng.integers(0, 8100, size=(P, n, 128)).astype(np.int16)
    rv = rng.integers(0, 8, size=(P, n, 128)).astype(np.int32)
    rv[:, :, 0] = 1
    rb = rng.integers(-7, 8, size=(P, n, 90)).astype(np.int8)
    rp = np.tile(np.array([1, -1] * (P // 2), np.int8)[:, None], (1, n))
    rr = rng.normal(size=(P, n))
    lens = rng.integers(1, P + 1, size=n)
    lens[:8] = [P, P, 61, 60, 59, 31, 30, 29]          # both sides of every length threshold
    played = np.arange(P)[:, None] < lens[None, :]
    winner = rng.choice([2, 0, 1, -1], size=n).astype(np.int8)
    winner[:8] = [1, -1, 1, -1, 0, 1, -1, 0]
    reason = rng.integers(1, 9, size=n).astype(np.uint8)
    player = rng.choice([1, -1], size=n).astype(np.int8)
    return rb, rp, rm, rv, rn, rr, played, winner, reason, player, lens.astype(np.int32)


def _eager(args, g, red_only, T):
    rb, rp, rm, rv, rn, rr, played, winner = args[:8]
    game_data, step_rewards = [], []
    for p in range(played.shape[0]):
        if not played[p, g]:
            break
        k = int(rn[p, g])
        counts = rv[p, g, :k].astype(np.int64)
        if T < 0.01:
            probs = np.zeros(k)
            probs[np.argmax(counts)] = 1
        else:
            c = counts ** (1.0 / T)
            probs = c / c.sum()
        pl = int(rp[p, g])
        if pl == 1 or not red_only:
            game_data.append((rb[p, g].reshape(10, 9).copy(),
                              {sp.unpack_move(m): pr for m, pr in zip(rm[p, g, :k].tolist(), probs)}, pl))
        step_rewards.append(float(rr[p, g]))
    w = int(winner[g])
    w = 0 if w == 2 else w
    return sp._shape_rewards(game_data, step_rewards, w), w


def test_materialise_arrays_equals_per_sample_formulation():
    args = _random_batch(np.random.default_rng(0))
    for red_only in (False, True):
        for T in (1.0, 0.5, 0.3, 0.0):
            res = sp.materialise_arrays(*args, T, red_only)
            assert len(res) == args[0].shape[1]
            for g, (data, winner, reason) in enumerate(res):
                want, w = _eager(args, g, red_only, T)
                assert winner == w and isinstance(reason, str) and len(data) == len(want)
                for (b1, d1, r1), (b2, d2, r2) in zip(want, data):
                    assert np.array_equal(b1, b2) and b2.shape == (10, 9) and b2.dtype == np.int8
                    assert type(r2) is float and repr(r1) == repr(r2)
                    assert list(d1.items()) == list(d2.items())
                    assert all(type(v) is np.float64 for v in d2.values())


def test_lazy_move_probs_behaves_like_the_dict_it_stands_for():
    args = _random_batch(np.random.default_rng(1), n=8)
    lazy = sp.materialise_arrays(*args, 1.0)[0][0][0][1]
    eager = sp.materialise_arrays(*args, 1.0, lazy=False)[0][0][0][1]
    assert lazy == eager and eager == lazy and len(lazy) == len(eager) > 0 and bool(lazy)
    assert list(lazy) == list(eager) and next(iter(lazy)) in lazy
    back = pickle.loads(pickle.dumps(lazy))
    assert type(back) is dict and back == dict(eager)
    assert sp._LazyMoveProbs() == {} and not sp._LazyMoveProbs()


def test_device_final_reward_is_float64_exact_on_every_branch():
    """samples._final_reward (torch, float64) == self_play.final_reward (the reference's branches,
    self_play.py:268-298) bit for bit, including 1.0 + 0.3 and -1.2, which float32 constants would
    not reproduce."""
    cases = [(w, p, n) for w in (0, 1, -1) for p in (1, -1) for n in (1, 29, 30, 31, 49, 50, 51, 59, 60, 61, 70, 71)]
    w = torch.tensor([c[0] for c in cases])
    p = torch.tensor([c[1] for c in cases])
    n = torch.tensor([c[2] for c in cases])
    got = _final_reward(w, p, n)
    assert got.dtype == torch.float64
    want = [sp.final_reward(*c) for c in cases]
    assert [repr(float(x)) for x in got] == [repr(x) for x in want]

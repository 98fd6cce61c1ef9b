"""Where the UNMODIFIED reference checkout (hpy666666/ChineseChessAI) lives, for tests and the
reference arm of bench.py.  Never imported by the product package.

Resolution order: ``$XQ_REFERENCE`` -> ``baseline/_ref`` (a git-ignored copy staged by
``scripts/stage_reference.py`` / ``__graft_entry__.build()``; it travels to the GPU box with the
gpurun snapshot) -> ``/root/reference`` (the authoring container only).
"""
from __future__ import annotations

import os
from typing import Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(ROOT, "integration", "shims")
_NEEDED = ("chess_env.py", "self_play.py", "neural_network.py", "config.py", "trainer.py")


def locate() -> Optional[str]:
    """Directory of the reference checkout, or None if it is not available here."""
    for cand in (os.environ.get("XQ_REFERENCE"), STAGED, "/root/reference"):
        if cand and all(os.path.isfile(os.path.join(cand, f)) for f in _NEEDED):
            return cand
    return None


def env_for_reference(hide_cuda: bool = True, threads: Optional[int] = None) -> dict:
    """Environment for a subprocess that runs the stock reference (its modules first on the
    path, nothing of this repo importable by accident, CUDA hidden so config.DEVICE == "cpu")."""
    ref = locate()
    if ref is None:
        raise FileNotFoundError("reference checkout not found (XQ_REFERENCE / baseline/_ref / /root/reference)")
    env = dict(os.environ)
    env["PYTHONPATH"] = ref
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    env["PYTHONIOENCODING"] = "utf-8"
    if hide_cuda:
        env["CUDA_VISIBLE_DEVICES"] = ""
    if threads is not None:
        env["OMP_NUM_THREADS"] = env["MKL_NUM_THREADS"] = str(max(1, int(threads)))
    return env


def env_for_shims() -> dict:
    """Environment for a subprocess that runs the reference's UNCHANGED consumers
    (trainer.py, evaluate.py, compare_models.py ...) on top of the three drop-in shim modules of
    INTEGRATION.md Option A: the shim directory shadows chess_env / self_play / neural_network,
    everything else (config.py included) comes from the reference checkout."""
    ref = locate()
    if ref is None:
        raise FileNotFoundError("reference checkout not found (XQ_REFERENCE / baseline/_ref / /root/reference)")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([SHIMS, ROOT, ref])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    env["PYTHONIOENCODING"] = "utf-8"
    return env

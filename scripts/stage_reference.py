#!/usr/bin/env python
"""Stage the unmodified reference's Python modules into the git-ignored ``baseline/_ref/`` so that
they travel to the GPU box with the gpurun snapshot (``/root/reference`` does not exist there).

The reference has no setup.py / pyproject.toml, so ``pip install --target baseline/_ref`` has
nothing to build; a verbatim copy of its top-level ``*.py`` files is the install.  The copy is
test/bench infrastructure only (never imported by the product package, never committed).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stage(src: str = "/root/reference", dst: str = os.path.join(ROOT, "baseline", "_ref")) -> int:
    if not os.path.isdir(src):
        return 0
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(src)):
        if not name.endswith(".py"):
            continue
        s, d = os.path.join(src, name), os.path.join(dst, name)
        data = open(s, "rb").read()
        if not (os.path.exists(d) and open(d, "rb").read() == data):
            shutil.copyfile(s, d)
        manifest[name] = hashlib.sha256(data).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1, sort_keys=True)
    return len(manifest)


if __name__ == "__main__":
    n = stage(*(sys.argv[1:3]))
    print(f"staged {n} reference modules into baseline/_ref" if n else "no reference checkout to stage")

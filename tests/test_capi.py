"""CPU: the C-ABI library builds, loads and exports exactly what include/xq_b200.h declares.
No compute entry point is called here (there is no GPU in this container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "xq_b200.h"), encoding="utf-8").read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(xq_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(built_lib):
    names = _declared()
    assert "xq_legal_moves" in names and "xq_step" in names and "xq_playout" in names
    raw = ctypes.CDLL(os.path.join(ROOT, "chinesechessai_b200", "libxq_b200.so"))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in xq_b200.h but not exported"


def test_binding_table_matches_header(built_lib):
    from chinesechessai_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_abi_version_and_meta_layout(built_lib):
    from chinesechessai_b200 import _lib
    assert built_lib.xq_abi_version() == _lib.ABI_VERSION
    assert _lib.META_DTYPE.itemsize == 32 and _lib.PLAYOUT_RESULT_DTYPE.itemsize == 40
    assert _lib.META_DTYPE.fields["move_count"][1] == 8
    assert _lib.META_DTYPE.fields["check_len"][1] == 28


def test_fails_loudly_without_device(built_lib):
    import pytest
    import torch
    from chinesechessai_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    assert built_lib.xq_device_count() < 0
    with pytest.raises(_lib.XqError):
        _lib.require_device()
    from chinesechessai_b200.engine import BoardBatch
    with pytest.raises(_lib.XqError):
        BoardBatch(4)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "chinesechessai_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", ""), \
                    f"{f} references oracle/"


def test_bench_reference_arm_prints_one_json_line():
    """bench.py's contract: exactly ONE line on stdout, a JSON object with the contract's keys.
    The reference arm (CPU port of the path) runs without a GPU, so it is checked here."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0", "--boards", "2048"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:300]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: on a machine without a CUDA device the product arm of bench.py exits
    with an error instead of timing something else."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--fast", "--no-cpu"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]

"""Build the host mirror of the device rules logic (test harness only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libxq_host_mirror.so")


def build() -> str:
    src = os.path.join(HERE, "xq_host_mirror.cpp")
    csrc = os.path.join(HERE, "..", "..", "chinesechessai_b200", "csrc")
    deps = [src, os.path.join(csrc, "xq_rules.cuh"), os.path.join(csrc, "xq_leap_table.inc"),
            os.path.join(csrc, "xq_knight_table.inc"), os.path.join(csrc, "xq_ray_table.inc")]
    if not os.path.exists(SO) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(SO):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared",
                               "-ffp-contract=off", "-Wall", "-o", SO, src])
    return SO

"""Builds the native library (csrc/*.cu -> csrc/librk.so) for sm_100a with nvcc."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "librk.so")
SOURCES = ["rk_interp.cu", "rk_optim.cu", "rk_drp.cu", "rk_tcgemm.cu", "rk_mlp2.cu", "rk_wgrad.cu", "rk_comm.cu", "rk_utility.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # no FMA contraction: keep per-op roundings like the reference
    "-Xcompiler", "-fPIC", "-shared", "-ldl",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".h", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "rk.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_native(force: bool = False, verbose: bool = False) -> str:
    from .core import isa
    isa.write_header(os.path.join(CSRC, "rk_isa.h")) if _header_changed() else None
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []),
           *[os.path.join(CSRC, s) for s in SOURCES], "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def _header_changed() -> bool:
    from .core import isa
    import io
    path = os.path.join(CSRC, "rk_isa.h")
    tmp = path + ".tmp"
    isa.write_header(tmp)
    new = open(tmp).read()
    os.remove(tmp)
    old = open(path).read() if os.path.exists(path) else ""
    return new != old


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Builds libpwc_b200.so (sm_100a only) in-tree with nvcc.  No GPU needed: nvcc cross-compiles."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_DIR = os.path.join(PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libpwc_b200.so")
SOURCES = ["pwc_abi.cu"]
HEADERS = ["common.cuh", "fwd_generic.cuh", "fwd_tile.cuh", "fwd_stream.cuh", "bwd.cuh", "bwd_fused.cuh", "fwd_mma16.cuh",
           os.path.join("..", "..", "include", "pwc_b200.h")]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for f in SOURCES + HEADERS + ["build.py"]:
        p = os.path.join(HERE, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False, out: str | None = None, extra_defs=()) -> str:
    if out is None and not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    lib_path = out or LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    defs = []
    if os.path.exists(os.path.join(HERE, "fwd_stream.cuh")):
        defs.append("-DPWC_HAVE_STREAM_KERNEL")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "--use_fast_math" if False else "-DPWC_NO_FAST_MATH", "-Xcompiler", "-fPIC,-O2", "-shared",
           "-cudart", "static", "-o", lib_path] + defs + list(extra_defs)
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(HERE, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libpwc_b200.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

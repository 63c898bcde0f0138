"""In-tree build of the CUDA extension (libzbphys.so) with nvcc for sm_100a."""

from __future__ import annotations

import os
import shutil
import subprocess

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = ["zbphys.cu"]
HEADERS = ["zb_kernel.cuh", "zb_math.cuh", os.path.join("..", "..", "include", "zbphys.h")]
LIB = os.path.join(_CSRC, "libzbphys.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(_CSRC, f)) > t for f in SOURCES + HEADERS)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> csrc/libzbphys.so (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB, *[os.path.join(_CSRC, s) for s in SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB

"""In-tree build of the CUDA extension (libzbphys.so) with nvcc for sm_100a.

``build_extension()`` builds the product library.  ``build_variant(name, defines)`` builds tuning / mutation variants of the
same sources into ``csrc/variants/libzbphys_<name>.so`` (``-DZB_MINIMAL``: only the plain step and reset instantiations, a
quarter of the compile time); the engine loads one with ``ZbotCudaEngine(..., lib_path=...)`` or ``ZBPHYS_LIB``.
"""

from __future__ import annotations

import os
import shutil
import subprocess

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = ["zbphys.cu"]
HEADERS = ["zb_kernel.cuh", "zb_math.cuh", os.path.join("..", "..", "include", "zbphys.h")]
LIB = os.path.join(_CSRC, "libzbphys.so")
VARIANT_DIR = os.path.join(_CSRC, "variants")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]

# Mutation builds (tests/test_parity_tail.py): each plants ONE bug in a rarely taken path of the step kernel; the parity
# checks must turn red on every one of them ("make the parity tests able to fail on a tail bug").
MUTANTS = {
    "mut1": ["-DZB_MUTANT=1"],  # wrong sign on the second tangent of manifold slot 2 (only feet with three contact points are affected)
    "mut2": ["-DZB_MUTANT=2"],  # the line search's full re-summation hands the wrong active set to a swapped bracket end
    "mut3": ["-DZB_MUTANT=3"],  # joint-limit rows drop out of the line search's zone test
    "mut4": ["-DZB_MUTANT=4"],  # one-sided loop (production instantiations only): the moved bracket end written back to the wrong side
}


def _newer_than_sources(path: str) -> bool:
    if not os.path.exists(path):
        return False
    t = os.path.getmtime(path)
    return not any(os.path.getmtime(os.path.join(_CSRC, f)) > t for f in SOURCES + HEADERS)


def _nvcc_cmd(out: str, defines=(), verbose: bool = False) -> list[str]:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return [nvcc, *NVCC_FLAGS, *defines, *(["-Xptxas", "-v"] if verbose else []), "-o", out, *[os.path.join(_CSRC, s) for s in SOURCES]]


def _build_two_units(out: str, defines=(), verbose: bool = False) -> None:
    """The full library = two translation units of zbphys.cu compiled in parallel (ZB_TU = 0: host code + the kernels of plane-box-only
    models, ZB_TU = 1: the kernels of the general collision groups), then linked."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(_CSRC, SOURCES[0])
    objs, procs = [], []
    for tu in (0, 1):
        obj = out[:-3] + f".tu{tu}.o"
        objs.append(obj)
        cmd = [nvcc, *[f for f in NVCC_FLAGS if f != "-shared"], "-c", f"-DZB_TU={tu}", *defines, *(["-Xptxas", "-v"] if verbose else []), "-o", obj, src]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate()[0] for p in procs]
    if any(p.returncode != 0 for p in procs):
        raise RuntimeError("nvcc failed:\n" + "\n".join(logs))
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    for o in objs:
        os.remove(o)
    if verbose:
        print("\n".join(logs))


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> csrc/libzbphys.so (nvcc cross-compiles without a GPU)."""
    if not force and _newer_than_sources(LIB):
        return LIB
    _build_two_units(LIB, (), verbose)
    return LIB


def variant_path(name: str) -> str:
    return os.path.join(VARIANT_DIR, f"libzbphys_{name}.so")


def build_variants(variants: dict[str, list[str]], force: bool = False, minimal: bool = True) -> dict[str, str]:
    """Build several variants concurrently (one nvcc process each)."""
    os.makedirs(VARIANT_DIR, exist_ok=True)
    procs, out = [], {}
    for name, defines in variants.items():
        path = variant_path(name)
        out[name] = path
        if not force and _newer_than_sources(path):
            continue
        if not minimal:  # a full two-unit build (sequential across variants, parallel inside)
            _build_two_units(path, tuple(defines))
            continue
        d = list(defines) + ["-DZB_MINIMAL"]
        procs.append((name, subprocess.Popen(_nvcc_cmd(path, d), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for variant {name}:\n{log}")
    return out


def build_mutants(force: bool = False) -> dict[str, str]:
    return build_variants(MUTANTS, force=force, minimal=True)

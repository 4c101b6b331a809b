"""Builds libb200ls.so (hand-written sm_100a CUDA + the C ABI of include/b200ls.h) in-tree with nvcc.

nvcc cross-compiles without a GPU; the built .so sits next to this file so that it travels to the
GPU box with the repository snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200ls.so")
SOURCES = ["b200ls.cu", "repart.cpp", "staggered.cpp"]
DEPS = ["b200ls.cu", "repart.cpp", "staggered.cpp", "sep_kernels.cuh", "sep_tile.cuh", "sep_solver.inc", "mg_kernels.cuh", "mg_schedule.h", "mg_solver.inc", "ops_kernels.cuh", "ops_solver.inc", "dense_kernels.cuh", "dense_solver.inc", "kernels.cuh", "spmv2.cuh", "spmv3.cuh", "spmv4.cuh", "spmv5.cuh", "update_fly.cuh", "hw.cuh", "csr_kernels.cuh", "csr_solver.inc", os.path.join("..", "..", "include", "b200ls.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",                      # no implicit FMA contraction: operand order of the reference is kept
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default",
    "-shared",
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200ls.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

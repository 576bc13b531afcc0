"""Build recipe for csrc/libwot_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libwot_b200.so")
SOURCES = ["api.cu", "solver.cu", "cost.cu", "pca.cu"]
HEADERS = ["common.cuh", "solver_state.cuh", "fused_iter.cuh", "fused_cluster.cuh", "online_pass.cuh", "online_tc.cuh", "online_solve.cuh", os.path.join("..", "..", "include", "wot_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > built for f in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    """Compile the CUDA library in-tree.  nvcc cross-compiles without a GPU."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))

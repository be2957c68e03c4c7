"""Builds libblokus_b200.so in-tree with nvcc for sm_100a (no other architecture, no JIT cache)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libblokus_b200.so")
SOURCES = ["blokus_b200.cu"]
DEPS = ["blokus_b200.cu", "movegen_core.cuh", "orient_gen.cuh", "tables_gen.h", "mcts.cuh", "heuristic.cuh",
        os.path.join("..", "..", "include", "blokus_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libblokus_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for d in DEPS:
        p = os.path.join(CSRC, d)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


PROBE_LIB = os.path.join(HERE, "libbk_probe.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libblokus_b200.so (and the bench-only pipe-throughput probe) if stale."""
    if force or needs_build():
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
        proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
        if verbose:
            sys.stderr.write(proc.stderr)
    probe_src = os.path.join(CSRC, "probe.cu")
    if force or not os.path.exists(PROBE_LIB) or os.path.getmtime(PROBE_LIB) < os.path.getmtime(probe_src):
        proc = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-o", PROBE_LIB, "probe.cu"], cwd=CSRC,
                              capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed (probe):\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

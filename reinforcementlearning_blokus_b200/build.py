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
DEPS = ["blokus_b200.cu", "movegen_core.cuh", "orient_gen.cuh", "tables_gen.h", "mcts.cuh", "mcts_wide.cuh", "heuristic.cuh",
        "snapshot.cuh", "probe.cu", os.path.join("..", "..", "include", "blokus_b200.h")]
PROBE_LIB = os.path.join(HERE, "libbk_probe.so")
STAMP = LIB + ".srchash"          # hash of the sources the library was built from (mtimes do not survive a copy)

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


def _source_hash() -> str:
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in DEPS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


PLAYOUT_KERNEL_SOURCES = ["blokus_b200.cu", "movegen_core.cuh", "orient_gen.cuh", "tables_gen.h"]


def playout_source_hash() -> str:
    """Hash of the files k_playout<0,false> is compiled from: profiles/*_playout_traffic.json records the hash its
    instruction / DRAM-byte constants were measured with, and bench.py flags them as stale when it differs."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in PLAYOUT_KERNEL_SOURCES:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def needs_build() -> bool:
    """True when the library is missing or was built from other sources (content hash, not mtimes: the tree is
    copied to the GPU box and a copy does not keep modification times in order)."""
    if not (os.path.exists(LIB) and os.path.exists(PROBE_LIB) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()




def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libblokus_b200.so and the bench-only pipe-throughput probe if missing or stale."""
    if force or needs_build():
        for out, srcs in ((LIB, SOURCES), (PROBE_LIB, ["probe.cu"])):
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + srcs
            proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            if proc.returncode != 0:
                raise RuntimeError(f"nvcc failed ({srcs[0]}):\n" + proc.stdout + proc.stderr)
            if verbose:
                sys.stderr.write(proc.stderr)
        with open(STAMP, "w") as f:
            f.write(_source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

#!/usr/bin/env python
"""Build recipe for ``oracle/_ref/``: the reference's own Python hot path as a compiled, importable artefact.

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  The reference is pure Python, so "compiling it from the
sources where they lie" means byte-compiling: every module listed below is compiled with ``py_compile`` straight from
``/root/reference`` and stored source-less in ONE importable archive, ``oracle/_ref/pyref.zip`` (zipimport loads
``.pyc`` members; loose ``*.pyc`` files do not survive the transfer to the GPU box).  ``oracle/_ref/`` is git-ignored,
NOT gpurun-ignored, so it travels to the GPU box exactly like a built ``.so``.  No reference source file is copied into
the repository.

Who may use the result (and only as checker / baseline, never on the product path):
  * ``tests/test_gpu_reference_suite.py`` -- runs the reference's OWN hot-path tests with the device generator and
    device agents installed (``integration.install(replace_class=True)`` + ``install_agents()``);
  * ``bench.py`` -- ``cpu_baseline`` / ``--impl reference`` time the Python engine (``kind: "reference"``).

The interpreter here and on the GPU box is the same image (CPython 3.12); the manifest records the bytecode magic
and ``ref_path()`` refuses to hand out a tree built by another interpreter.
"""

from __future__ import annotations

import glob
import hashlib
import importlib.util
import json
import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BLOKUS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
TREE = os.path.join(OUT, "pyref.zip")
MANIFEST = os.path.join(OUT, "MANIFEST.json")

# The hot path of SURVEY.md section 8(a) and its callers, the benchmark of BASELINE configs[0] and the reference's own
# hot-path tests.  Globs are relative to the reference root.
MODULES = [
    "engine/*.py",
    "mcts/*.py",
    "agents/__init__.py", "agents/random_agent.py", "agents/heuristic_agent.py", "agents/fast_mcts_agent.py",
    "agents/gameplay_fast_mcts.py",
    "analytics/winprob/__init__.py", "analytics/winprob/features.py", "analytics/winprob/dataset.py",
    "analytics/tournament/__init__.py", "analytics/tournament/arena_runner.py", "analytics/tournament/arena_stats.py",
    "benchmarks/benchmark_move_generation.py",
    "tests/__init__.py", "tests/utils_game_states.py",
]
# the reference's hot-path test files (collected through stub modules, see tests/test_gpu_reference_suite.py)
TEST_MODULES = [
    "tests/test_move_generation_equivalence.py", "tests/test_legality_bitboard_equivalence.py",
    "tests/test_bitboard_basic.py", "tests/test_frontier_basic.py", "tests/test_pieces_orientations.py",
    "tests/test_piece_shapes_match.py", "tests/test_engine.py", "tests/test_game_over_logic.py",
    "tests/test_game_result.py", "tests/test_fast_mcts_think.py", "tests/test_mcts_diagnostics.py",
    "tests/test_telemetry.py",
]
# (tests/test_mobility_metrics.py is left out: it opens a JSON fixture next to its own source file, which an archive
#  of byte-code does not have; the formula it checks is pinned by tests/golden/mobility.npz instead)


def _files():
    out = []
    for pat in MODULES + TEST_MODULES:
        hits = sorted(glob.glob(os.path.join(REF, pat)))
        if not hits:
            raise FileNotFoundError(f"reference file missing: {pat}")
        out.extend(os.path.relpath(h, REF) for h in hits)
    return out


def build(force: bool = False) -> str:
    """Byte-compile the listed reference modules into oracle/_ref/pyref.zip.  Needs /root/reference (build container)."""
    if not os.path.isdir(os.path.join(REF, "engine")):
        raise FileNotFoundError(f"no reference checkout at {REF}: oracle/_ref can only be built in the build container")
    files = _files()
    digest = {}
    for rel in files:
        with open(os.path.join(REF, rel), "rb") as f:
            digest[rel] = hashlib.sha256(f.read()).hexdigest()
    magic = importlib.util.MAGIC_NUMBER.hex()
    if not force and os.path.exists(MANIFEST):
        with open(MANIFEST) as f:
            old = json.load(f)
        if old.get("sha256") == digest and old.get("magic") == magic:
            return TREE
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(TREE, "w", zipfile.ZIP_DEFLATED) as z:
        dirs = set()
        for rel in files:                              # explicit directory entries: zipimport needs them to see the
            parts = rel.split("/")[:-1]                # packages without __init__ (analytics, benchmarks) as namespaces
            for i in range(1, len(parts) + 1):
                dirs.add("/".join(parts[:i]) + "/")
        for d in sorted(dirs):
            z.writestr(zipfile.ZipInfo(d), b"")
        for rel in files:
            dst = os.path.join(tmp, "m.pyc")
            py_compile.compile(os.path.join(REF, rel), cfile=dst, dfile="reference:/" + rel, doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            z.write(dst, rel[:-3] + ".pyc")
    with open(MANIFEST, "w") as f:
        json.dump({"what": "byte-compiled modules of the reference (TGALLOWAY1/ReinforcementLearning_Blokus), built by "
                           "oracle/make_ref.py from " + REF, "python": sys.version.split()[0], "magic": magic,
                   "tests": [t[:-3].replace("/", ".") for t in TEST_MODULES], "sha256": digest}, f, indent=1)
    return TREE


def ref_path() -> str:
    """Path to put on sys.path to import the reference (engine, mcts, agents, analytics, benchmarks, tests); raises if
    the artefact is missing or was built by another interpreter."""
    if not os.path.exists(MANIFEST):
        raise FileNotFoundError("oracle/_ref is missing: run `python oracle/make_ref.py` in the build container "
                                "(__graft_entry__.build() does)")
    with open(MANIFEST) as f:
        m = json.load(f)
    if m.get("magic") != importlib.util.MAGIC_NUMBER.hex():
        raise RuntimeError(f"oracle/_ref was built by Python {m.get('python')}; this is {sys.version.split()[0]}")
    if not os.path.isfile(TREE):
        raise FileNotFoundError("oracle/_ref/pyref.zip is missing (manifest present): rebuild with oracle/make_ref.py")
    return TREE


def available() -> bool:
    try:
        ref_path()
        return True
    except (FileNotFoundError, RuntimeError):
        return False


def test_modules():
    with open(MANIFEST) as f:
        return list(json.load(f)["tests"])


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, and exports exactly
the entry points include/blokus_b200.h declares.  No compute calls (no GPU here)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from reinforcementlearning_blokus_b200 import build

    return build.build()


def _header_functions():
    with open(os.path.join(ROOT, "include", "blokus_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/blokus_b200.h but not exported"


def test_python_prototypes_cover_header(lib_path):
    from reinforcementlearning_blokus_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == _header_functions()
    lib = _lib.load()
    assert b"sm_100a" in lib.bk_version()


def test_state_struct_is_256_bytes():
    from reinforcementlearning_blokus_b200.state import STATE_DTYPE, empty_states

    assert STATE_DTYPE.itemsize == 256
    assert STATE_DTYPE.fields["used"][1] == 224 and STATE_DTYPE.fields["to_move"][1] == 240
    s = empty_states(3)
    assert s["first_mask"].tolist() == [15, 15, 15]


def test_compute_fails_loudly_without_gpu(lib_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from reinforcementlearning_blokus_b200 import _lib, device

    with pytest.raises(_lib.BlokusB200Error):
        device.playout(4, seed=1)
    # and the raw ABI refuses compute before bk_init
    lib = _lib.load()
    assert lib.bk_movegen_count(None, None, 1, None, None) == -3
    assert b"bk_init" in lib.bk_last_error()


def test_sass_is_sm100a_only(lib_path):
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs

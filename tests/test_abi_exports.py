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


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/blokus_b200.h compiles as C11 (the boundary is a C ABI) and bk_mcts_params / bk_state have the byte
    layout the ctypes binding assumes."""
    import ctypes
    import shutil
    import subprocess

    from reinforcementlearning_blokus_b200._lib import MctsParams

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    fields = [f[0] for f in MctsParams._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "blokus_b200.h"\nint main(void){\n'
                   + 'printf("%zu %zu\\n", sizeof(bk_mcts_params), sizeof(bk_state));\n'
                   + "".join(f'printf("{f} %zu\\n", offsetof(bk_mcts_params, {f}));\n' for f in fields) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call([gcc, "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    lines = subprocess.check_output([str(exe)], text=True).split("\n")
    size_params, size_state = (int(x) for x in lines[0].split())
    assert size_params == ctypes.sizeof(MctsParams) and size_state == 256
    for line in lines[1:]:
        if line:
            name, off = line.split()
            assert getattr(MctsParams, name).offset == int(off), name


def test_init_failure_raises_and_does_not_deadlock(lib_path):
    """a failing bk_init used to re-enter the module lock.  Without a GPU (or with a device index that does
    not exist) init() must raise BlokusB200Error, twice in a row."""
    import threading

    from reinforcementlearning_blokus_b200 import _lib

    result = []

    def run():
        for _ in range(2):
            try:
                _lib.init(9999)
                result.append("no error")
            except _lib.BlokusB200Error as exc:
                result.append(str(exc))

    t = threading.Thread(target=run, daemon=True)
    t.start()
    t.join(timeout=60)
    assert not t.is_alive(), "init() hung"
    assert len(result) == 2 and all("blokus_b200 error" in r for r in result), result

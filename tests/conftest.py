import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(1, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """Build the CUDA library (and the oracle) if this tree has not been built yet -- nvcc cross-compiles without a
    GPU, so this works on both boxes.  Building is not a fallback: without the library every product call raises."""
    try:
        from reinforcementlearning_blokus_b200 import build

        if build.needs_build():
            build.build()
    except Exception as exc:  # report, and let the tests that need the library fail loudly
        sys.stderr.write(f"[conftest] could not build libblokus_b200.so: {exc}\n")
    try:  # the byte-compiled reference (checker / CPU baseline); only the build container has the sources
        from oracle import make_ref

        if os.path.isdir(os.path.join(make_ref.REF, "engine")):
            make_ref.build()
    except Exception as exc:
        sys.stderr.write(f"[conftest] could not build oracle/_ref: {exc}\n")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN

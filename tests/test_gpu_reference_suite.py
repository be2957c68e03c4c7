"""The reference's own hot-path tests (BASELINE.json configs[2] names
tests/test_legality_bitboard_equivalence.py:206-318) judged against the DEVICE path on a B200.

oracle/_ref holds the reference's modules and test modules, byte-compiled by oracle/make_ref.py (build container).
A child pytest imports them, installs the device generator as ``engine.move_generator.LegalMoveGenerator`` and as the
shared generator, ``Board.can_place_piece`` on the device, and the device ``MCTSAgent`` / ``FastMCTSAgent`` adapters,
and runs the reference's tests unchanged.  In those tests the reference's naive generator and grid rules stay the
oracle (``_get_legal_moves_naive`` is not replaced), the device answers for ``get_legal_moves`` /
``_get_legal_moves_frontier`` / ``is_placement_legal_bitboard_coords`` / ``has_legal_moves`` / ``is_move_legal`` /
``can_place_piece`` and ``FastMCTSAgent.think``.
"""

import json
import os

import pytest

import ref_suite

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_tests_pass_with_the_device_path_installed(tmp_path):
    from oracle import make_ref

    assert torch.cuda.is_available()
    modules = make_ref.test_modules()        # raises when oracle/_ref is missing: build() makes it, it ships with the tree
    res = ref_suite.run(modules, str(tmp_path / "ref"), device=True, extra_args=["-rA"])
    tail = "\n".join(res["output"].strip().splitlines()[-60:])
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "reference_suite.json"), "w") as f:
            json.dump({k: v for k, v in res.items() if k != "output"} | {"modules": modules, "tail": tail}, f, indent=1)
    # what the child process had installed when it collected the reference's tests
    assert res["report"] == ("generator class: B200LegalMoveGenerator; shared: B200LegalMoveGenerator; FastMCTSAgent: "
                             "B200FastMCTSAgent; MCTSAgent: B200MCTSAgent; Board.can_place_piece: _device_can_place_piece"), res["report"] + tail
    assert res["returncode"] == 0 and res["failed"] == 0 and res["error"] == 0, tail
    assert res["passed"] >= 105, tail

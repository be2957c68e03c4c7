"""oracle/_ref (the byte-compiled reference, oracle/make_ref.py) imports and runs on its own -- no device involved.
Skipped where the artefact was never built (a checkout without the build container's /root/reference)."""

import pytest

import ref_suite
from oracle import make_ref

pytestmark = pytest.mark.skipif(not make_ref.available(), reason="oracle/_ref not built (needs the reference checkout)")


def test_manifest_lists_the_hot_path_tests():
    mods = make_ref.test_modules()
    assert "tests.test_legality_bitboard_equivalence" in mods and "tests.test_fast_mcts_think" in mods
    assert len(mods) >= 12


def test_reference_tests_run_from_the_artifact(tmp_path):
    res = ref_suite.run(["tests.test_pieces_orientations", "tests.test_piece_shapes_match", "tests.test_game_result"],
                        str(tmp_path / "ref"), device=False)
    assert res["returncode"] == 0 and res["passed"] >= 25, res["output"][-2000:]


def test_every_module_of_the_artifact_imports(tmp_path):
    """Namespace packages (analytics, benchmarks) included: the search agents import analytics.winprob."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r); sys.dont_write_bytecode = True\n"
            "import engine.game, mcts.mcts_agent, agents.fast_mcts_agent, analytics.tournament.arena_runner\n"
            "import benchmarks.benchmark_move_generation, tests.utils_game_states\n"
            "print(engine.game.__file__)") % make_ref.ref_path()
    out = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "pyref.zip" in out.stdout, out.stderr[-2000:]

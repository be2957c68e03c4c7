"""The reference-side shim (integration.py).  Runs only where the reference checkout exists (the build
container); on the GPU box it is skipped -- nothing there may read /root/reference."""

import json
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("BLOKUS_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "engine")), reason="reference checkout not present")


@pytest.fixture()
def reference_on_path():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    yield
    sys.path.remove(REF)
    for m in [m for m in sys.modules if m == "engine" or m.startswith("engine.")]:
        del sys.modules[m]


def test_adapter_packs_and_installs(reference_on_path, golden_dir):
    from engine.board import Board, Player, Position
    from engine.move_generator import Move
    import engine.move_generator as mg

    from reinforcementlearning_blokus_b200 import integration
    from reinforcementlearning_blokus_b200.board import Board as MyBoard

    with open(os.path.join(golden_dir, "kat2_game.json")) as f:
        kat = json.load(f)
    board = Board()
    mine = MyBoard()
    gen = mg.LegalMoveGenerator()
    from reinforcementlearning_blokus_b200 import pieces as P
    for mv in kat["moves"]:
        pid, ori, r, c = P.decode_move(mv["move_id"])
        m = Move(pid, ori, r, c)
        pl = Player(mv["player"] + 1)
        assert board.place_piece(m.get_positions(gen.piece_orientations_cache[pid]), pl, pid)
        from reinforcementlearning_blokus_b200.board import Player as MyPlayer, Position as MyPosition
        cells = [MyPosition(r + dr, c + dc) for dr, dc in P.PIECE_ORIENTATIONS[pid][ori].cells]
        assert mine.place_piece(cells, MyPlayer(mv["player"] + 1), pid, validate=False)
    packed = integration.pack_reference_board(board)
    assert packed.tobytes().hex() == kat["packed_hex"]            # reference object -> bk_state
    assert mine.to_packed().tobytes().hex() == kat["packed_hex"]  # host mirror -> the same bytes
    adapter = integration.make_adapter()
    assert isinstance(adapter, mg.LegalMoveGenerator)
    assert set(adapter.piece_orientations_cache) == set(range(1, 22))
    try:
        installed = integration.install()
        assert mg.get_shared_generator() is installed
    finally:
        integration.uninstall()
    assert mg._SHARED_GENERATOR is not installed

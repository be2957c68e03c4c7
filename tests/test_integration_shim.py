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


def test_agent_adapters_patch_and_convert(reference_on_path, monkeypatch):
    """install_agents() swaps MCTSAgent / FastMCTSAgent wherever the reference holds them; the adapters take the
    reference's objects and hand back the reference's Move (the device search itself is stubbed here: no GPU)."""
    import agents.fast_mcts_agent as ref_fast
    import mcts.mcts_agent as ref_mcts
    from engine.board import Board, Player
    from engine.move_generator import Move

    from reinforcementlearning_blokus_b200 import agents as my_agents
    from reinforcementlearning_blokus_b200 import integration
    from reinforcementlearning_blokus_b200 import mcts as my_mcts
    seen = {}

    def fake_select(self, board, player, legal_moves):
        seen["board"], seen["player"], seen["n"] = board, player, len(legal_moves)
        return legal_moves[1]

    def fake_think(self, board, player, legal_moves, time_budget_ms):
        return {"move": legal_moves[0], "stats": {"timeBudgetMs": time_budget_ms, "nodesEvaluated": 7}}

    monkeypatch.setattr(my_mcts.MCTSAgent, "select_action", fake_select)
    monkeypatch.setattr(my_agents.FastMCTSAgent, "think", fake_think)
    orig_mcts, orig_fast = ref_mcts.MCTSAgent, ref_fast.FastMCTSAgent
    try:
        new_mcts, new_fast = integration.install_agents()
        assert ref_mcts.MCTSAgent is new_mcts and ref_fast.FastMCTSAgent is new_fast
        board = Board()
        legal = [Move(1, 0, 0, 0), Move(2, 0, 0, 0), Move(2, 1, 0, 0)]
        agent = ref_mcts.MCTSAgent(iterations=10, seed=1, use_transposition_table=False)
        got = agent.select_action(board, Player.RED, legal)
        assert isinstance(got, Move) and (got.piece_id, got.orientation) == (2, 0)
        from reinforcementlearning_blokus_b200.board import Board as MyBoard

        assert isinstance(seen["board"], MyBoard) and len(seen["board"].to_packed().tobytes()) == 256
        assert seen["player"].value == 1 and seen["n"] == 3
        out = ref_fast.FastMCTSAgent(iterations=5, time_limit=0.01, seed=2).think(board, Player.RED, legal, 25)
        assert isinstance(out["move"], Move) and out["stats"]["timeBudgetMs"] == 25
    finally:
        integration.uninstall_agents()
    assert ref_mcts.MCTSAgent is orig_mcts and ref_fast.FastMCTSAgent is orig_fast

"""The object API on a GPU, written the way the reference tests its own engine (tests/test_engine.py,
tests/test_legality_bitboard_equivalence.py, tests/test_game_over_logic.py, tests/test_game_result.py,
tests/test_fast_mcts_think.py, tests/test_mcts_diagnostics.py) plus parity against the pinned oracle."""

import math
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402
from reinforcementlearning_blokus_b200 import (BlokusGame, Board, FastMCTSAgent, GameplayFastMCTSAgent,  # noqa: E402
                                               HeuristicAgent, LegalMoveGenerator, Move, Player, Position,
                                               RandomAgent, get_shared_generator)


# ---- tests/test_engine.py:94-153 ---------------------------------------------------------------------
def test_first_move_corner_rule():
    board = Board()
    assert board.can_place_piece([Position(0, 0), Position(0, 1)], Player.RED)
    assert not board.can_place_piece([Position(1, 1), Position(1, 2)], Player.RED)
    assert board.can_place_piece([Position(0, 19)], Player.BLUE)
    assert board.can_place_piece([Position(19, 19), Position(18, 19)], Player.YELLOW)
    assert board.can_place_piece([Position(19, 0)], Player.GREEN)
    assert not board.can_place_piece([], Player.RED)
    assert not board.can_place_piece([Position(-1, 0)], Player.RED)


def test_edge_and_corner_adjacency_rules():
    board = Board()
    assert board.place_piece([Position(0, 0), Position(0, 1)], Player.RED, 1)
    assert board.can_place_piece([Position(1, 2), Position(1, 3)], Player.RED)          # corner contact
    assert not board.can_place_piece([Position(0, 2), Position(0, 3)], Player.RED)      # edge contact
    assert not board.can_place_piece([Position(5, 5)], Player.RED)                      # no contact
    assert not board.can_place_piece([Position(0, 1)], Player.RED)                      # occupied
    board.place_piece([Position(1, 2), Position(1, 3)], Player.RED, 2)
    assert board.can_place_piece([Position(2, 4), Position(2, 5)], Player.RED)
    board.assert_bitboard_consistent()
    assert board.current_player == Player.YELLOW and board.move_count == 2


def test_generator_interface_and_first_moves():
    gen = LegalMoveGenerator()
    assert len(gen.all_pieces) == 21 and len(gen.piece_orientations_cache) == 21
    board = Board()
    for p in Player:
        moves = gen.get_legal_moves(board, p)
        assert len(moves) == 58
        corner = board.player_start_corners[p]
        for m in moves[::7]:
            assert gen.is_move_legal(board, p, m)
            assert corner in m.get_positions(gen.piece_orientations_cache[m.piece_id])
    assert not gen.is_move_legal(board, Player.RED, Move(1, 0, -1, -1))      # tests/test_engine.py:260-274
    assert not gen.is_move_legal(board, Player.RED, Move(99, 0, 0, 0))
    assert not gen.is_move_legal(board, Player.RED, Move(1, 5, 0, 0))
    assert gen.has_legal_moves(board, Player.RED) and gen.get_move_count(board, Player.RED) == 58
    assert get_shared_generator() is get_shared_generator()


def test_game_flow_matches_oracle_move_by_move():
    """A BlokusGame driven by RandomAgent; the oracle replays the same moves and must agree on every
    legal-move list, score and on the game-over moment (engine/game.py:57-214)."""
    game = BlokusGame(enable_telemetry=False)
    agent = RandomAgent(seed=11)
    ob = orc.OracleBoard()
    assert game.board.current_player == Player.RED and game.move_count == 0 and not game.is_game_over()
    turns = 0
    while not game.is_game_over() and turns < 120:
        p = game.get_current_player()
        legal = game.get_legal_moves(p)
        assert [m.to_id() for m in legal] == ob.legal_moves(p.value - 1).tolist()
        turns += 1
        if not legal:
            game.board._update_current_player()
            game._check_game_over()
            ob.pass_turn()
            ob.check_game_over()
            continue
        move = agent.select_action(game.board, p, legal)
        assert game.make_move(move, p)
        assert ob.make_move(p.value - 1, move.to_id())
        assert game.is_game_over() == ob.game_over
        if turns % 9 == 0:
            for q in Player:
                assert game.get_score(q) == ob.game_score(q.value - 1)
                assert game.board.get_score(q) == ob.board_score(q.value - 1)
    assert game.is_game_over() and ob.game_over
    res = game.get_game_result()
    assert res.scores == {q.value: ob.game_score(q.value - 1) for q in Player}
    top = max(res.scores.values())
    assert res.winner_ids == [k for k, v in res.scores.items() if v == top] and res.is_tie == (len(res.winner_ids) > 1)
    assert (game.winner is None) == res.is_tie
    assert game.board.to_packed().tobytes() == ob.pack().tobytes()
    assert len(game.game_history) == game.move_count
    # rebuilt from the packed state, the board is the same board
    b2 = Board.from_packed(game.board.to_packed())
    assert np.array_equal(b2.grid, game.board.grid) and b2.player_pieces_used == game.board.player_pieces_used


def test_invalid_moves_and_piece_exhaustion():
    game = BlokusGame()
    assert not game.make_move(Move(1, 0, -1, -1))
    assert game.move_count == 0
    for i in range(1, 22):
        game.board.player_pieces_used[Player.RED].add(i)
    assert game.get_legal_moves(Player.RED) == [] and not game.can_player_move(Player.RED)
    assert game.board.get_score(Player.RED) == 15                         # all-pieces bonus, board.py:574
    game.reset_game()
    assert game.move_count == 0 and game.board.current_player == Player.RED and game.winner is None
    first = game.get_legal_moves()[0]
    assert game.make_move(first) and game.move_count == 1 and game.board.current_player == Player.BLUE
    assert not game.make_move(first, Player.RED)                          # piece already spent
    st = game.get_game_state()
    assert {"board", "current_player", "move_count", "game_over", "scores"} <= set(st)


def test_frontier_is_exact_after_moves():
    """tests/test_frontier_basic.py:107-164: start corner first, then diagonal cells of the placed piece."""
    board = Board()
    assert board.get_frontier(Player.RED) == {(0, 0)}
    board.place_piece([Position(0, 0)], Player.RED, 1)
    assert board.get_frontier(Player.RED) == {(1, 1)}
    board.place_piece([Position(0, 19), Position(0, 18)], Player.BLUE, 2)
    assert board.get_frontier(Player.BLUE) == {(1, 17)}


def test_heuristic_scores_match_reference(golden_dir):
    """agents/heuristic_agent.py:68-103 scores from the reference itself (tests/golden/heuristic.npz)."""
    z = np.load(os.path.join(golden_dir, "heuristic.npz"))
    states = z["states"].view(orc.STATE_DTYPE).reshape(-1)
    offs = z["offsets"]
    agent = HeuristicAgent(seed=0)
    for i, s in enumerate(states):
        board = Board.from_packed(s)
        p = Player(int(s["to_move"]) + 1)
        moves = [Move.from_id(int(m)) for m in z["move_ids"][offs[i]:offs[i + 1]]]
        got = agent.evaluate_moves(board, p, moves)
        assert np.array_equal(got, z["scores"][offs[i]:offs[i + 1]]), i          # float64, bit for bit
        pick = agent.select_action(board, p, moves)
        assert pick in moves
    assert agent.select_action(Board(), Player.RED, []) is None


def test_fast_mcts_think_schema_and_budget():
    """tests/test_fast_mcts_think.py and tests/test_mcts_diagnostics.py of the reference."""
    board = Board()
    gen = get_shared_generator()
    legal = gen.get_legal_moves(board, Player.RED)
    agent = FastMCTSAgent(iterations=5000, time_limit=1.0, seed=7)
    import time
    t0 = time.perf_counter()
    out = agent.think(board, Player.RED, legal, 150)
    elapsed = (time.perf_counter() - t0) * 1000
    assert elapsed < 1500
    assert out["move"] in legal
    st = out["stats"]
    for k in ("timeBudgetMs", "timeSpentMs", "nodesEvaluated", "maxDepthReached", "topMoves"):
        assert k in st
    assert st["timeBudgetMs"] == 150 and st["nodesEvaluated"] >= 5 and len(st["topMoves"]) <= 10
    for tm in st["topMoves"]:
        assert {"piece_id", "orientation", "anchor_row", "anchor_col", "visits", "q_value"} <= set(tm)
    # diagnostics invariants
    agent = FastMCTSAgent(iterations=300, time_limit=60.0, seed=7, enable_diagnostics=True, diagnostics_sample_interval=100)
    out = agent.think(board, Player.RED, legal, 60000)
    d = out["stats"]["diagnostics"]
    assert d["version"] == "v1" and d["simulations"] == 300 and d["rootLegalMoves"] == 58
    assert sum(x["nodes"] for x in d["nodesByDepth"]) == d["nodesExpanded"] + 1
    assert 0.0 <= d["policyEntropy"] <= math.log(max(d["rootChildrenExpanded"], 1)) + 1e-9
    assert [t["sim"] for t in d["bestMoveTrace"]] == [100, 200]
    assert sum(tm["visits"] for tm in d["rootPolicy"]) <= 300
    # same seed, same search
    a1 = FastMCTSAgent(iterations=200, time_limit=60.0, seed=3).think(board, Player.RED, legal, 60000)
    a2 = FastMCTSAgent(iterations=200, time_limit=60.0, seed=3).think(board, Player.RED, legal, 60000)
    assert a1["move"] == a2["move"] and a1["stats"]["topMoves"] == a2["stats"]["topMoves"]
    # degenerate inputs
    assert agent.think(board, Player.RED, [], 10)["move"] is None
    one = agent.think(board, Player.RED, legal[:1], 10)
    assert one["move"] == legal[0] and one["stats"]["nodesEvaluated"] == 1
    mv, stats = GameplayFastMCTSAgent(iterations=64, seed=1).choose_move(board, Player.RED, legal, 200)
    assert mv in legal and "topMoves" in stats


def test_heuristic_agent_custom_weights():
    """HeuristicAgent.set_weights (agents/heuristic_agent.py:268-282; the arena's "weights" parameter): the kernel takes
    the four weights and combines the same features in the reference's order -- checked against the default-weight
    scores recombined on the host, and through select_action."""
    from reinforcementlearning_blokus_b200 import BlokusGame, HeuristicAgent, RandomAgent
    from reinforcementlearning_blokus_b200 import device as dev

    game = BlokusGame()
    rnd = RandomAgent(seed=3)
    for _ in range(14):
        p = game.get_current_player()
        game.make_move(rnd.select_action(game.board, p, game.get_legal_moves(p)), p)
    board, player = game.board, game.get_current_player()
    legal = game.get_legal_moves(player)
    base = HeuristicAgent(seed=1)
    s_default = base.evaluate_moves(board, player, legal)
    # isolate each feature with unit weights, then recombine with arbitrary ones
    parts = []
    for k in range(4):
        a = HeuristicAgent(seed=1)
        w = [0.0, 0.0, 0.0, 0.0]; w[k] = 1.0
        a.set_weights(dict(zip(("piece_size", "corner_creation", "edge_avoidance", "center_preference"), w)))
        parts.append(a.evaluate_moves(board, player, legal))
    assert np.array_equal(((0.0 + 1.0 * parts[0]) + 2.0 * parts[1] + -1.5 * parts[2]) + 0.5 * parts[3], s_default)
    custom = HeuristicAgent(seed=1)
    custom.set_weights({"piece_size": 0.7, "corner_creation": 3.25, "edge_avoidance": -0.4, "center_preference": 1.9})
    got = custom.evaluate_moves(board, player, legal)
    want = (((0.0 + 0.7 * parts[0]) + 3.25 * parts[1]) + -0.4 * parts[2]) + 1.9 * parts[3]
    assert np.array_equal(got, want) and not np.array_equal(got, s_default)
    assert custom.select_action(board, player, legal) in legal
    assert custom.get_action_info()["weights"]["corner_creation"] == 3.25


# ---- the C ABI keeps one context per device -------------------------------------------------------------------------------
def test_bad_device_raises_instead_of_hanging():
    from reinforcementlearning_blokus_b200 import _lib

    with pytest.raises(_lib.BlokusB200Error):
        _lib.init(9999)
    with pytest.raises(_lib.BlokusB200Error):          # a second failure must not dead-lock either
        _lib.init(9999)


def test_tensor_on_another_device_is_refused():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from reinforcementlearning_blokus_b200 import device as dev
    from reinforcementlearning_blokus_b200.state import empty_states

    with torch.cuda.device(1):
        s1 = dev.to_device(empty_states(2))
    with torch.cuda.device(0), pytest.raises(ValueError):
        dev.movegen_count(s1)


def test_two_devices_in_one_process():
    """bk_init(1) after bk_init(0) must not disturb device 0 (per-device contexts; the current device is left alone)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from reinforcementlearning_blokus_b200 import device as dev

    outs = {}
    for d in (0, 1, 0, 1):
        with torch.cuda.device(d):
            r = dev.playout(256, seed=5, want_trace=True)
            got = (r["game_scores"].cpu(), r["trace"].cpu())
            assert torch.cuda.current_device() == d
        if d in outs:
            assert torch.equal(outs[d][0], got[0]) and torch.equal(outs[d][1], got[1])
        outs[d] = got
    assert torch.equal(outs[0][0], outs[1][0])


def test_peer_root_merge_single_rank():
    """bk_root_merge_p2p with one rank: the merged statistics are the sums over the rank's own trees (the slots alternate
    and are cleared after every merge, so repeated merges of different sizes stay exact)."""
    from reinforcementlearning_blokus_b200 import dist as bdist

    peer = bdist.PeerRootMerge(capacity=2048)
    g = torch.Generator(device="cuda").manual_seed(3)
    for it in range(12):
        n_trees, n = 1 + it % 4, (1, 58, 237, 2048)[it % 4]
        vis = torch.randint(0, 1 << 20, (n_trees, 2048), generator=g, device="cuda", dtype=torch.int32)
        rew = torch.randint(-(1 << 40), 1 << 40, (n_trees, 2048), generator=g, device="cuda", dtype=torch.int64)
        v, w = peer.merge(vis, rew, n)
        assert torch.equal(v, vis.sum(dim=0, dtype=torch.int64)[:n]) and torch.equal(w, rew.sum(dim=0)[:n])
    # many trees: torch adds them up first and the kernel takes int64 visit sums
    vis = torch.randint(0, 1 << 20, (100, 2048), generator=g, device="cuda", dtype=torch.int32)
    rew = torch.randint(-(1 << 40), 1 << 40, (100, 2048), generator=g, device="cuda", dtype=torch.int64)
    v, w = peer.merge(vis, rew, 237)
    assert torch.equal(v, vis.sum(dim=0, dtype=torch.int64)[:237]) and torch.equal(w, rew.sum(dim=0)[:237])
    with pytest.raises(ValueError):
        peer.merge(vis.to(torch.int16), rew, 5)
    with pytest.raises(ValueError):
        peer.merge(vis, rew, 4096)
    peer.close()


def test_peer_root_merge_two_ranks():
    """Two processes, two GPUs (torchrun, NCCL for the comparison): 200 merges over NVLink peer memory equal the all-reduce
    on every rank (tools/p2p_merge_check.py exits non-zero otherwise)."""
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", "29631", "tools/p2p_merge_check.py"], cwd=root, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]

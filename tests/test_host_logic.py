"""Host-side logic that needs no GPU: board bookkeeping, packing, move ids, frontier view, arena utilities.

Literal cases follow the reference's own unit tests (tests/test_bitboard_basic.py:21-40,137-153,
tests/test_frontier_basic.py:107-164, tests/test_pieces_orientations.py, tests/test_engine.py:21-92)."""

import numpy as np
import pytest

from reinforcementlearning_blokus_b200 import pieces as P
from reinforcementlearning_blokus_b200.board import Board, Player, Position
from reinforcementlearning_blokus_b200.move_generator import LegalMoveGenerator, Move
from reinforcementlearning_blokus_b200.state import STATE_DTYPE, bits_to_int, empty_states, grid_of, int_to_bits


def test_board_initial_state_and_cells():
    b = Board()
    assert b.grid.shape == (20, 20) and not b.grid.any()
    assert b.current_player == Player.RED and b.move_count == 0 and not b.game_over
    assert b.player_start_corners[Player.RED] == Position(0, 0)
    assert b.player_start_corners[Player.BLUE] == Position(0, 19)
    assert b.player_start_corners[Player.YELLOW] == Position(19, 19)
    assert b.player_start_corners[Player.GREEN] == Position(19, 0)
    assert b.is_valid_position(Position(19, 19)) and not b.is_valid_position(Position(20, 0))
    assert b.is_empty(Position(5, 5)) and b.get_cell(Position(-1, 0)) == -1
    b.set_cell(Position(5, 5), 2)
    assert b.get_player_at(Position(5, 5)) == Player.BLUE and b.get_player_at(Position(0, 0)) is None
    assert len(b.get_adjacent_positions(Position(0, 0))) == 3
    assert len(b.get_edge_adjacent_positions(Position(10, 10))) == 4
    assert len(b.get_corner_adjacent_positions(Position(0, 19))) == 1


def test_bit_layout_literals():
    """bit = row*20 + col (engine/bitboard.py:13-22); two dominoes give the reference's literal masks."""
    b = Board()
    b.place_piece([Position(0, 0), Position(0, 1)], Player.RED, 2, validate=False)
    b.place_piece([Position(0, 19), Position(1, 19)], Player.BLUE, 2, validate=False)
    assert b.player_bits[Player.RED] == (1 << 0) | (1 << 1)
    assert b.player_bits[Player.BLUE] == (1 << 19) | (1 << 39)
    assert b.occupied_bits == b.player_bits[Player.RED] | b.player_bits[Player.BLUE]
    b.assert_bitboard_consistent()
    s = b.to_packed()
    assert int(s["bits"][0][0]) == 3 and int(s["bits"][1][0]) == (1 << 19) | (1 << 39)
    assert int(s["used"][0]) == 2 and int(s["first_mask"]) == 0b1100 and int(s["to_move"]) == 2 and int(s["move_count"]) == 2
    assert bits_to_int(int_to_bits(b.occupied_bits)) == b.occupied_bits
    assert np.array_equal(grid_of(s), b.grid)


def test_pack_roundtrip_and_copy():
    rng = np.random.RandomState(0)
    b = Board()
    for k in range(12):
        pl = b.current_player
        cells = [Position(int(rng.randint(20)), int(rng.randint(20)))]
        if b.grid[cells[0].row, cells[0].col] == 0:
            b.place_piece(cells, pl, 1 + k, validate=False)
        else:
            b._update_current_player()
    s = b.to_packed()
    assert s.dtype == STATE_DTYPE and s.tobytes() == Board.from_packed(s).to_packed().tobytes()
    c = b.copy()
    c.place_piece([Position(10, 10)], c.current_player, 21, validate=False)
    assert c.move_count == b.move_count + 1 and not np.array_equal(c.grid, b.grid)
    assert b.to_packed().tobytes() == s.tobytes()          # the copy is deep
    e = empty_states(1)[0]
    assert Board().to_packed().tobytes() == e.tobytes()


def test_frontier_literals():
    """tests/test_frontier_basic.py:107-164 of the reference."""
    b = Board()
    for p in Player:
        c = b.player_start_corners[p]
        assert b.get_frontier(p) == {(c.row, c.col)}
    b.place_piece([Position(0, 0)], Player.RED, 1, validate=False)
    assert b.get_frontier(Player.RED) == {(1, 1)}
    b2 = Board()
    b2.place_piece([Position(0, 0), Position(0, 1)], Player.RED, 2, validate=False)
    assert b2.get_frontier(Player.RED) == {(1, 2)}
    # a cell taken by an opponent stays in the owner's incremental set (engine/board.py:315-367 touches the mover's
    # set only); the exact view drops it.  Literals from the reference run on the same placements.
    b2.place_piece([Position(1, 2)], Player.BLUE, 1, validate=False)
    assert b2.get_frontier(Player.RED) == {(1, 2)} and b2._compute_full_frontier(Player.RED) == set()
    assert b2.get_frontier(Player.BLUE) == {(2, 3), (0, 19), (0, 3), (2, 1)}
    assert b2._compute_full_frontier(Player.BLUE) == {(0, 3), (2, 1), (2, 3)}
    assert b2.copy().get_frontier(Player.RED) == {(1, 2)}


def test_move_ids_cover_every_orientation():
    gen = LegalMoveGenerator()
    assert len(gen.all_pieces) == 21 and sum(p.size for p in gen.all_pieces) == 88
    seen = set()
    for o in P.ORIENTATIONS:
        m = Move(o.piece_id, o.index, 20 - o.height, 20 - o.width)
        mid = m.to_id()
        assert mid == P.encode_move(o.piece_id, o.index, 20 - o.height, 20 - o.width)
        assert Move.from_id(mid) == m and mid not in seen
        seen.add(mid)
        assert Move(o.piece_id, o.index, 21 - o.height, 0).to_id() is None      # hangs over the bottom edge
        cells = [(p.row, p.col) for p in m.get_positions(gen.piece_orientations_cache[o.piece_id])]
        assert cells == [(20 - o.height + r, 20 - o.width + c) for r, c in o.cells]
        assert gen.piece_position_cache[o.piece_id][o.index] == list(o.cells)
    assert Move(1, 0, -1, -1).to_id() is None and Move(22, 0, 0, 0).to_id() is None and Move(1, 1, 0, 0).to_id() is None
    assert max(seen) < P.MAX_MOVE_ID and P.decode_move(0) == (1, 0, 0, 0)


def test_arena_utilities_without_gpu():
    from reinforcementlearning_blokus_b200 import arena

    assert arena.compute_ranks({"1": 10, "2": 10, "3": 3, "4": 7}) == {"1": 1, "2": 1, "3": 3, "4": 2}
    sa = arena.seat_assignment_for_game(["a", "b", "c", "d"], 2, 0, "round_robin")
    assert sa == {"1": "c", "2": "d", "3": "a", "4": "b"}
    rec = arena._record("run", 3, 99, sa, "round_robin", {"1": 5, "2": 9, "3": 9, "4": 1}, 40, 50, 10, 0, 0.5, False, {}, None)
    assert rec["winner_ids"] == [2, 3] and rec["is_tie"] and rec["winner_id"] is None
    assert rec["winner_agents"] == ["d", "a"] and rec["agent_scores"]["c"] == 5 and rec["game_id"] == "run_g0003"


def test_arena_run_config_parsing():
    """RunConfig / AgentConfig / SnapshotConfig: the reference's keys, defaults, legacy shape and validation errors
    (analytics/tournament/arena_runner.py:52-236); the literal below is what the reference's own parser returns."""
    import pytest

    from reinforcementlearning_blokus_b200 import arena

    cfg = {"agents": [{"name": "a", "type": "random"}, {"name": "b", "type": "heuristic", "thinking_time_ms": 50},
                      {"name": "c", "type": "fast_mcts", "iterations": 200}, {"name": "d", "type": "mcts", "params": {"iterations": 64}}],
           "num_games": 3, "seed": 11, "seat_policy": "round_robin", "snapshots": {"enabled": True, "checkpoints": [16, 8, 8]}}
    got = arena.RunConfig.from_dict(cfg).to_dict()
    assert got == {
        "agents": [{"name": "a", "type": "random", "thinking_time_ms": None, "params": {}},
                   {"name": "b", "type": "heuristic", "thinking_time_ms": 50, "params": {}},
                   {"name": "c", "type": "fast_mcts", "thinking_time_ms": None, "params": {"iterations": 200}},
                   {"name": "d", "type": "mcts", "thinking_time_ms": None, "params": {"iterations": 64}}],
        "num_games": 3, "seed": 11, "seat_policy": "round_robin", "output_root": "arena_runs", "max_turns": 2500, "notes": "",
        "snapshots": {"enabled": True, "strategy": "fixed_ply", "checkpoints": [8, 16]}}
    legacy = {"x": {"type": "random"}, "y": {"type": "heuristic", "time_limit": 0.2}, "z": {"type": "random"},
              "w": {"type": "random"}, "num_games": 2, "seed": 1}
    lg = arena.RunConfig.from_dict(legacy)
    assert lg.seat_policy == "randomized" and lg.agents[1].thinking_time_ms == 200 and lg.agents[1].params == {"time_limit": 0.2}
    for bad in ({**cfg, "num_games": 0}, {**cfg, "agents": cfg["agents"][:3]}, {**cfg, "seat_policy": "fixed"},
                {**cfg, "agents": [dict(a, name="same") for a in cfg["agents"]]}, {**cfg, "snapshots": {"strategy": "x"}}):
        with pytest.raises(ValueError):
            arena.RunConfig.from_dict(bad)
    with pytest.raises(ValueError):
        arena.build_agent(arena.AgentConfig("q", "alphazero"), seed=1)


def test_pairwise_dataset_matches_reference(golden_dir):
    """winprob.build_pairwise_dataset / split_pairwise_by_game against the reference's output on the same frame
    (tests/golden/pairwise.npz, written by generate_golden.py pairwise): same rows, order, columns and floats."""
    import importlib.util
    import json
    import os

    from reinforcementlearning_blokus_b200 import winprob

    spec = importlib.util.spec_from_file_location("generate_golden_input", os.path.join(golden_dir, "generate_golden.py"))
    z = np.load(os.path.join(golden_dir, "pairwise.npz"))
    # rebuild the input frame with the generator's recipe without importing the reference
    src = open(spec.origin).read()
    start = src.index("def pairwise_input_frame():")
    ns = {"np": np, "os": os, "json": json, "HERE": golden_dir}
    exec(src[start:src.index("def gen_pairwise():")], ns)
    df, cols = ns["pairwise_input_frame"]()
    out, meta = winprob.build_pairwise_dataset(df)
    assert list(out.columns) == z["columns"].tolist() and meta == json.loads(str(z["meta"][0]))
    assert np.array_equal(out[cols].to_numpy(dtype=np.float64), z["X"]) and np.array_equal(out["label"].to_numpy(), z["label"])
    assert np.array_equal(out["player_i_id"].to_numpy(), z["player_i"]) and np.array_equal(out["player_j_id"].to_numpy(), z["player_j"])
    assert out["game_id"].tolist() == z["game"].tolist() and out["agent_i"].tolist() == z["agent_i"].tolist()
    assert out["agent_j"].tolist() == z["agent_j"].tolist() and out["phase_bucket"].tolist() == z["bucket"].tolist()
    assert np.array_equal(out["phase_board_occupancy"].to_numpy(), z["occupancy"])
    train, test, split = winprob.split_pairwise_by_game(out, test_size=0.25, seed=3)
    assert split == json.loads(str(z["meta"][1]))
    assert np.array_equal(train.index.to_numpy(), z["train_index"]) and np.array_equal(test.index.to_numpy(), z["test_index"])


def test_device_model_forms_match_sklearn(golden_dir, tmp_path):
    """The two model forms handed to the UCT kernel -- folded logistic weights and flattened gradient-boosted phase
    models -- evaluated by the CPU oracle (the kernel's restatement) agree with scikit-learn's predict_proba used the way
    mcts/learned_evaluator.py:97-143 uses it, on reference feature rows."""
    import os

    import pytest

    joblib = pytest.importorskip("joblib")
    from sklearn.ensemble import GradientBoostingClassifier
    from sklearn.linear_model import LogisticRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    from oracle import orc
    from reinforcementlearning_blokus_b200 import winprob

    z = np.load(os.path.join(golden_dir, "snapshots.npz"))
    rows = z["state_features"]                                                   # [240, 4, 34]
    states = np.load(os.path.join(golden_dir, "random_states.npz"))["states"].view(orc.STATE_DTYPE).reshape(-1)
    cols = winprob.SNAPSHOT_FEATURE_COLUMNS
    x = np.concatenate([rows[:, 0] - rows[:, 1], rows[:, 2] - rows[:, 3], rows[:, 1] - rows[:, 2]])
    y = (x[:, cols.index("mobility_total_placements")] + 8000 * x[:, cols.index("player_board_occupancy")] > 0).astype(int)
    logreg = Pipeline([("scale", StandardScaler()), ("clf", LogisticRegression(max_iter=500))]).fit(x, y)
    occ = np.concatenate([(rows[:, a, 29] + rows[:, b, 29]) / 2 for a, b in ((0, 1), (2, 3), (1, 2))])
    early = GradientBoostingClassifier(n_estimators=40, max_depth=3, random_state=1).fit(x[occ < 0.4], y[occ < 0.4])
    late = GradientBoostingClassifier(n_estimators=25, max_depth=2, random_state=2).fit(x, y)
    art_l = {"model_type": "pairwise_logreg", "feature_columns": cols, "pipeline": logreg}
    art_g = {"model_type": "pairwise_gbt_phase", "feature_columns": cols, "phase_models": {"early": early}, "fallback_model": late}
    w, blob = winprob.fold_logistic_pipeline(art_l), winprob.flatten_gbt_phase_models(art_g)
    assert w.shape == (35,) and blob is not None and blob[1] == blob[2] != blob[0]           # mid and late share the fallback
    for i in (0, 17, 101, 230):
        for player in (0, 3):
            want_l, want_g = [], []
            for j in range(4):
                if j == player:
                    continue
                d = (rows[i, player] - rows[i, j])[None]
                want_l.append(logreg.predict_proba(d)[0, 1])
                o = (rows[i, player, 29] + rows[i, j, 29]) / 2.0
                want_g.append((early if o < 0.33 else late).predict_proba(d)[0, 1])
            assert orc.win_probability(states[i], player, w, 0) == pytest.approx(float(np.mean(want_l)), abs=1e-12)
            assert orc.win_probability(states[i], player, blob, 1) == pytest.approx(float(np.mean(want_g)), abs=1e-12)
    # artifacts of another shape are declined (the host evaluator still serves them out of tree)
    assert winprob.fold_logistic_pipeline(art_g) is None and winprob.flatten_gbt_phase_models(art_l) is None
    assert winprob.flatten_gbt_phase_models({"model_type": "pairwise_gbt_phase", "phase_models": {"early": early}}) is None
    p = str(tmp_path / "gbt.joblib")
    joblib.dump(art_g, p)
    ev = winprob.LearnedWinProbabilityEvaluator(p)
    assert ev.device_model_kind == 1 and np.array_equal(ev.device_model, blob)


def test_arena_budget_rule_and_summary():
    """Deterministic budget rule of the arena adapters (arena_runner.py:334-356, 446-450) and the per-agent summary."""
    from reinforcementlearning_blokus_b200 import arena

    cap = arena._iteration_cap
    assert cap(arena.AgentConfig("a", "gameplay_fast_mcts", 25, {"iterations_per_ms": 20.0})) == 500
    assert cap(arena.AgentConfig("a", "fast_mcts", 200)) == 4000                       # default 20 iterations per ms
    assert cap(arena.AgentConfig("a", "mcts", 30)) == 300                              # default 10 iterations per ms
    assert cap(arena.AgentConfig("a", "mcts", None, {"iterations": 64})) == 64
    assert cap(arena.AgentConfig("a", "fast_mcts", 100, {"deterministic_time_budget": False, "iterations": 500})) == 500
    recs = [
        {"error": None, "agent_scores": {"x": 70, "y": 60, "z": 60, "w": 10}, "agent_ranks": {"x": 1, "y": 2, "z": 2, "w": 3},
         "winner_agents": ["x"]},
        {"error": None, "agent_scores": {"x": 50, "y": 80, "z": 80, "w": 20}, "agent_ranks": {"x": 2, "y": 1, "z": 1, "w": 3},
         "winner_agents": ["y", "z"]},
        {"error": "boom", "agent_scores": {}, "agent_ranks": {}, "winner_agents": []},
    ]
    s = arena.summarize(recs, ["x", "y", "z", "w"])
    assert s["completed_games"] == 2 and s["error_games"] == 1
    assert s["agents"]["x"] == {"games": 2, "wins": 1.0, "win_rate": 0.5, "mean_score": 60.0, "mean_rank": 1.5}
    assert s["agents"]["y"]["wins"] == 0.5 and s["agents"]["w"]["win_rate"] == 0.0
    assert arena.compute_ranks({"1": 30, "2": 30, "3": 10, "4": 50}) == {"1": 2, "2": 2, "3": 3, "4": 1}       # dense ranks


def test_snapshot_rows_schema():
    """Rows of snapshots.csv: key order of arena_runner.py:_build_snapshot_rows_for_checkpoint (:552-575) followed by the
    34 feature columns, labels filled in from the finished game (:722-731); checkpoints a game never reached are skipped."""
    import torch

    from reinforcementlearning_blokus_b200 import winprob

    feats = torch.arange(2 * 2 * 4 * 34, dtype=torch.float64).reshape(2, 2, 4, 34)
    replay = {"features": feats, "turn_index": torch.tensor([[8, 17], [9, -1]], dtype=torch.int32), "checkpoints": np.array([8, 16])}
    recs = [{"run_id": "r", "game_id": "r_g0000", "game_index": 0, "game_seed": 5, "seat_assignment": {"1": "a", "2": "b", "3": "c", "4": "d"},
             "final_scores": {"1": 50, "2": 60, "3": 60, "4": 10}, "final_ranks": {"1": 2, "2": 1, "3": 1, "4": 3},
             "winner_ids": [2, 3], "winner_id": None, "is_tie": True},
            {"run_id": "r", "game_id": "r_g0001", "game_index": 1, "game_seed": 6, "seat_assignment": {"1": "b", "2": "c", "3": "d", "4": "a"},
             "final_scores": {"1": 90, "2": 60, "3": 61, "4": 10}, "final_ranks": {"1": 1, "2": 3, "3": 2, "4": 4},
             "winner_ids": [1], "winner_id": 1, "is_tie": False}]
    rows = winprob.snapshot_rows_from_replay(replay, recs)
    assert len(rows) == 4 * 3                                              # game 1 never reached ply 16
    assert list(rows[0])[:20] == ["run_id", "game_id", "game_index", "game_seed", "checkpoint_index", "checkpoint_ply", "ply",
                                  "turn_index", "current_player_id", "player_id", "player_name", "agent_name", "seat_index",
                                  "winner_id", "winner_ids_json", "label_is_winner", "final_score", "final_rank",
                                  "final_scores_json", "is_tie"]
    assert list(rows[0])[20:] == winprob.SNAPSHOT_FEATURE_COLUMNS
    r = rows[5]                                                            # game 0, checkpoint 1 (ply 16), player 2
    assert (r["checkpoint_index"], r["checkpoint_ply"], r["turn_index"], r["player_id"], r["player_name"], r["agent_name"]) == (1, 16, 17, 2, "BLUE", "b")
    assert r["current_player_id"] == 17 % 4 + 1 and r["label_is_winner"] == 1 and r["winner_ids_json"] == "[2, 3]" and r["is_tie"] is True
    assert r["frontier_size"] == float(feats[0, 1, 1, 0]) and rows[-1]["winner_id"] == 1 and rows[-1]["final_rank"] == 4


def test_move_list_forgets_its_order_when_changed():
    """get_legal_moves returns a list that knows it is in ascending id order (agents skip re-sorting it); any in-place
    change or copy makes it an ordinary list again."""
    import random

    from reinforcementlearning_blokus_b200.move_generator import MoveList

    def fresh():
        m = MoveList([10, 20, 30, 40])
        m.ascending_ids = np.array([10, 20, 30, 40])
        return m

    assert fresh().ascending_ids is not None
    assert getattr(fresh()[:2], "ascending_ids", None) is None and getattr(list(fresh()), "ascending_ids", None) is None
    for change in (lambda m: random.Random(0).shuffle(m), lambda m: m.sort(reverse=True), lambda m: m.pop(), lambda m: m.append(5),
                   lambda m: m.reverse(), lambda m: m.__setitem__(0, 1), lambda m: m.remove(20), lambda m: m.extend([1])):
        m = fresh()
        change(m)
        assert m.ascending_ids is None

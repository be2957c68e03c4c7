"""N2: the callers on top of the path -- arena match-ups (object loop and device batch) and the /think contract."""

import json

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402
from reinforcementlearning_blokus_b200 import (BlokusGame, FastMCTSAgent, HeuristicAgent, Player, RandomAgent,  # noqa: E402
                                               arena, service)
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.batched import BatchedBlokus  # noqa: E402


def _oracle_word(packed_i32):
    """seat-policy word for orc.playout: same bytes with bit 31 set (as a C int)."""
    return int(np.int32(packed_i32)) | -(1 << 31)


def test_mixed_seat_policies_match_oracle():
    n, seed = 64, 17
    rng = np.random.RandomState(0)
    seats = rng.randint(0, 2, size=(n, 4))
    packed = (seats[:, 0] | (seats[:, 1] << 8) | (seats[:, 2] << 16) | (seats[:, 3] << 24)).astype(np.int32)
    r = dev.playout(n, seed=seed, game_offset=9, seat_policies=torch.from_numpy(packed).cuda(), want_trace=True)
    tr, gs = r["trace"].cpu().numpy(), r["game_scores"].cpu().numpy()
    for i in range(0, n, 5):
        o = orc.playout(seed, 9 + i, policy=_oracle_word(packed[i]))
        assert np.array_equal(tr[i], o["trace"]), i
        assert np.array_equal(gs[i], o["game_scores"])
    # all-uniform and all-heuristic words reproduce the single-policy kernels
    for pol in (0, 1):
        word = torch.full((8,), pol * 0x01010101, dtype=torch.int32, device="cuda")
        a = dev.playout(8, seed=3, seat_policies=word, want_trace=True)["trace"]
        b = dev.playout(8, seed=3, policy=pol, want_trace=True)["trace"]
        assert torch.equal(a, b)


def test_seed_and_seat_utilities():
    assert arena.stable_hash_int(42, 0, "game_seed") == arena.game_seed_from_run_seed(42, 0)
    assert 0 <= arena.game_seed_from_run_seed(42, 7) < 2**31 - 1
    names = ["a", "b", "c", "d"]
    assert arena.seat_assignment_for_game(names, 1, 5, "round_robin") == {"1": "b", "2": "c", "3": "d", "4": "a"}
    sa = arena.seat_assignment_for_game(names, 3, 1234, "randomized")
    assert sorted(sa.values()) == names and sa == arena.seat_assignment_for_game(names, 3, 1234, "randomized")
    assert arena.compute_ranks({"1": 50, "2": 70, "3": 70, "4": 10}) == {"1": 2, "2": 1, "3": 1, "4": 3}   # dense, as the reference


def test_device_arena_heuristic_beats_random(tmp_path):
    out = arena.run_device_arena({"h1": "heuristic", "r1": "random", "h2": "heuristic", "r2": "random"}, 2048,
                                 run_seed=42, seat_policy="round_robin")
    recs, summ = out["records"], out["summary"]
    assert len(recs) == 2048 and sum(s["games"] for s in summ.values()) == 4 * 2048
    assert abs(sum(s["wins"] for s in summ.values()) - 2048) < 1e-6
    assert summ["h1"]["mean_score"] > summ["r1"]["mean_score"] + 3
    assert summ["h1"]["win_rate"] + summ["h2"]["win_rate"] > 0.6
    rec = recs[5]
    for key in ("run_id", "game_id", "game_index", "game_seed", "seat_assignment", "seat_policy", "winner_ids",
                "winner_agents", "is_tie", "final_scores", "final_ranks", "agent_scores", "agent_ranks", "winner_id",
                "moves_made", "turn_count", "passes", "invalid_actions", "duration_sec", "truncated", "agent_move_stats",
                "snapshot_checkpoints_hit", "error"):
        assert key in rec
    assert rec["game_seed"] == arena.game_seed_from_run_seed(42, 5)
    assert rec["seat_assignment"] == arena.seat_assignment_for_game(["h1", "r1", "h2", "r2"], 5, rec["game_seed"], "round_robin")
    # the record agrees with an oracle replay of the same game
    o = orc.playout(42, 5, policy=_oracle_word(out["seat_policies"][5]))
    assert rec["final_scores"] == {str(p + 1): int(o["game_scores"][p]) for p in range(4)}
    assert rec["moves_made"] == int((o["trace"] != 0xFFFF).sum())
    path = tmp_path / "games.jsonl"
    arena.write_games_jsonl(str(path), recs[:10])
    assert len([json.loads(line) for line in open(path)]) == 10


def test_object_arena_game_and_think_contract():
    agents = {"rand": RandomAgent(seed=1), "heur": HeuristicAgent(seed=2),
              "fast": FastMCTSAgent(iterations=40, time_limit=5.0, seed=3), "rand2": RandomAgent(seed=4)}
    seats = arena.seat_assignment_for_game(list(agents), 0, arena.game_seed_from_run_seed(1, 0), "round_robin")
    rec = arena.run_single_game(run_id="t", game_index=0, game_seed=arena.game_seed_from_run_seed(1, 0),
                                seat_assignment=seats, agents=agents, max_turns=12)
    assert rec["error"] is None and rec["truncated"] and rec["turn_count"] == 12 and rec["moves_made"] == 12
    assert rec["agent_move_stats"]["fast"]["moves"] == 3
    # POST /think contract (engine-service/app.py:53-72)
    game = BlokusGame()
    for _ in range(6):
        game.make_move(game.get_legal_moves()[0])
    legal = game.get_legal_moves()
    payload = {
        "gameState": {"board": game.board.grid.tolist(),
                      "pieces_used": {p.name: sorted(game.board.player_pieces_used[p]) for p in Player},
                      "current_player": game.board.current_player.name, "move_count": game.board.move_count},
        "legalMoves": [{"piece_id": m.piece_id, "orientation": m.orientation, "anchor_row": m.anchor_row,
                        "anchor_col": m.anchor_col} for m in legal],
        "timeBudgetMs": 300,
    }
    rebuilt = service.rebuild_board(payload["gameState"])
    assert rebuilt.to_packed().tobytes() == game.board.to_packed().tobytes()
    resp = service.think(payload, iterations=200)
    assert resp["move"] in payload["legalMoves"]
    assert {"timeBudgetMs", "timeSpentMs", "nodesEvaluated", "topMoves"} <= set(resp["stats"])
    assert service.think({"gameState": payload["gameState"], "legalMoves": [], "timeBudgetMs": 10})["move"] is None


def test_batched_blokus_api():
    b = BatchedBlokus(512)
    counts = b.legal_move_counts()
    assert counts.tolist() == [58] * 512
    _, offsets, ids = b.legal_moves()
    # everybody plays its first legal move (torch cannot index uint16 tensors: go through an int16 view)
    first = ids.view(torch.int16)[offsets[:-1].long()].view(torch.uint16)
    ok, alive = b.step(first, want_alive=True)
    assert ok.all() and (alive == 15).all() and (b.to_move() == 1).all()
    gs, bs = b.scores()
    assert (bs[:, 0] == 1).all() and (gs[:, 0] == 6).all()     # a monomino on the start corner: 1 cell + 5 corner bonus
    moved = b.step_uniform_random(seed=5, turn=1)
    assert (moved.cpu().numpy() != 0xFFFF).all() and (b.to_move() == 2).all()
    fin = b.playout(seed=5, first_turn=2, want_final=True)
    assert not dev.has_moves(fin["final_states"]).any()


def test_mobility_after_moves_matches_oracle():
    """N3: the movegen-bound inner loop of engine/telemetry.py:17-77 as one device batch."""
    from reinforcementlearning_blokus_b200.batched import mobility_after_moves

    o = orc.playout(5, 1, max_turns=14, want_final=True)
    state = o["final"]
    b = orc.OracleBoard.from_packed(state)
    player, opp = 0, 2
    legal = b.legal_moves(opp)
    sample = legal[:: max(1, len(legal) // 12)][:12]
    got = mobility_after_moves(state, player, opp, sample).cpu().numpy()
    for mv, g in zip(sample, got):
        c = b.copy()
        assert c.place(opp, int(mv))
        assert g == len(c.legal_moves(player))
    assert mobility_after_moves(state, player, opp, [36399]).cpu().numpy()[0] == -1


def test_playout_pipeline_matches_direct_calls():
    n = 2048
    from reinforcementlearning_blokus_b200.state import empty_states

    host = torch.from_numpy(empty_states(n).view(np.uint8).reshape(n, 256).copy()).pin_memory()
    pipe = dev.PlayoutPipeline(n)
    got = []
    for k in range(5):
        pipe.submit(host, 42, game_offset=k * n)
        r = pipe.collect()
        if r is not None:
            got.append(r.clone())
    got += pipe.drain()
    assert len(got) == 5
    for k in range(5):
        want = dev.playout(n, seed=42, game_offset=k * n)["game_scores"].cpu()
        assert torch.equal(got[k], want), k
    with pytest.raises(RuntimeError):
        pipe.submit(host, 1); pipe.submit(host, 1); pipe.submit(host, 1)
    pipe.drain()


def test_run_experiment_device_batch_and_object_agents(tmp_path):
    """arena_runner.run_experiment with the reference's config format: built-in policies run as one device batch,
    any other agent mix through the per-game loop; both write run_config.json, games.jsonl, snapshots.csv, summary.json."""
    import json
    import os

    from reinforcementlearning_blokus_b200 import arena

    base = {"num_games": 6, "seed": 5, "seat_policy": "randomized", "output_root": str(tmp_path / "runs"),
            "snapshots": {"enabled": True, "checkpoints": [8, 24, 40]}}
    dev_cfg = arena.RunConfig.from_dict({**base, "agents": [{"name": "r1", "type": "random"}, {"name": "h1", "type": "heuristic"},
                                                            {"name": "r2", "type": "random"}, {"name": "h2", "type": "heuristic"}]})
    out = arena.run_experiment(dev_cfg)
    run_dir = out["run_dir"]
    games = [json.loads(line) for line in open(os.path.join(run_dir, "games.jsonl"))]
    assert len(games) == 6 and out["snapshot_rows"] == 4 * sum(len(g["snapshot_checkpoints_hit"]) for g in games) > 0
    for gi, g in enumerate(games):            # seeds and seats exactly as the reference derives them
        assert g["game_seed"] == arena.game_seed_from_run_seed(5, gi)
        assert g["seat_assignment"] == arena.seat_assignment_for_game(["r1", "h1", "r2", "h2"], gi, g["game_seed"], "randomized")
    summary = json.load(open(os.path.join(run_dir, "summary.json")))
    assert summary["completed_games"] == 6 and abs(sum(a["wins"] for a in summary["agents"].values()) - 6) < 1e-9
    assert os.path.exists(os.path.join(run_dir, "snapshots.csv")) and os.path.exists(os.path.join(run_dir, "run_config.json"))
    # object agents: a leaf-batched FastMCTS agent with a deterministic budget against three random agents
    obj_cfg = arena.RunConfig.from_dict({**base, "num_games": 2, "agents": [
        {"name": "fast", "type": "fast_mcts", "thinking_time_ms": 5, "iterations_per_ms": 16, "leaf_batch": 8},
        {"name": "ra", "type": "random"}, {"name": "rb", "type": "random"}, {"name": "rc", "type": "random"}]})
    out2 = arena.run_experiment(obj_cfg)
    games2 = [json.loads(line) for line in open(os.path.join(out2["run_dir"], "games.jsonl"))]
    assert len(games2) == 2 and all(g["error"] is None and g["invalid_actions"] == 0 for g in games2)
    assert all("move_trace" not in g for g in games2) and out2["snapshot_rows"] > 0
    st = games2[0]["agent_move_stats"]["fast"]
    assert st["moves"] > 0 and st["avg_simulations_per_move"] == 80          # iteration cap = round(16 * 5)
    # the same built-in match-up through the object loop gives records of the same shape
    loop = arena.run_experiment(arena.RunConfig.from_dict({**base, "num_games": 1, "agents": dev_cfg.to_dict()["agents"]}),
                                device_batch=False)
    rec = json.loads(open(os.path.join(loop["run_dir"], "games.jsonl")).readline())
    assert set(rec) == set(games[0])


def test_device_arena_with_search_agents(tmp_path):
    """The shape of the reference's scripts/arena_config.json (four gameplay_fast_mcts agents with different budgets,
    round robin, snapshots) as device batches: legal complete games (replayed by the oracle), deterministic, the
    simulations per move follow the budget rule, and a mixed table of random / heuristic / mcts / fast_mcts agents."""
    import json
    import os

    from oracle import orc
    from reinforcementlearning_blokus_b200 import arena

    cfg = arena.RunConfig.from_dict({
        "agents": [{"name": f"mcts_{ms}ms", "type": "gameplay_fast_mcts", "thinking_time_ms": ms, "iterations_per_ms": 8}
                   for ms in (2, 4, 6, 8)],
        "num_games": 10, "seed": 20260301, "seat_policy": "round_robin", "output_root": str(tmp_path / "runs"),
        "snapshots": {"enabled": True, "checkpoints": [8, 32]}})
    a = arena.run_device_arena_agents(cfg)
    b = arena.run_device_arena_agents(cfg)
    assert torch.equal(a["trace"], b["trace"])
    one_warp = arena.run_device_arena_agents(cfg, auto_leaf_batch=False)         # one rollout per iteration, as the reference
    assert one_warp["summary"]["completed_games"] == 10
    assert {n: s["avg_simulations_per_move"] for n, s in one_warp["agent_move_stats"].items()} == {
        "mcts_2ms": 16.0, "mcts_4ms": 32.0, "mcts_6ms": 48.0, "mcts_8ms": 64.0}
    assert {n: s["avg_simulations_per_move"] for n, s in a["agent_move_stats"].items()} == {
        "mcts_2ms": 16.0, "mcts_4ms": 32.0, "mcts_6ms": 48.0, "mcts_8ms": 64.0}
    tr = a["trace"].cpu().numpy()
    for g, rec in enumerate(a["records"]):
        board = orc.OracleBoard()
        for t in range(rec["turn_count"]):
            if tr[g, t] == 0xFFFF:
                assert not board.has_legal_moves(board.current_player)
                board.pass_turn()
            else:
                assert board.make_move(board.current_player, int(tr[g, t]))
        assert board.check_game_over()
        assert {str(p + 1): board.game_score(p) for p in range(4)} == rec["final_scores"]
        assert rec["seat_assignment"] == arena.seat_assignment_for_game(cfg.agent_names, g, rec["game_seed"], "round_robin")
    assert len(a["snapshot_rows"]) == 4 * sum(len(r["snapshot_checkpoints_hit"]) for r in a["records"]) > 0
    # through run_experiment, with a mixed table
    mixed = arena.RunConfig.from_dict({
        "agents": [{"name": "rnd", "type": "random"}, {"name": "heur", "type": "heuristic"},
                   {"name": "uct", "type": "mcts", "thinking_time_ms": 3, "use_transposition_table": False},
                   {"name": "fast", "type": "fast_mcts", "thinking_time_ms": 2, "leaf_batch": 4}],
        "num_games": 8, "seed": 3, "seat_policy": "randomized", "output_root": str(tmp_path / "runs")})
    out = arena.run_experiment(mixed)
    games = [json.loads(line) for line in open(os.path.join(out["run_dir"], "games.jsonl"))]
    assert len(games) == 8 and out["summary"]["completed_games"] == 8
    assert games[0]["agent_move_stats"]["uct"]["avg_simulations_per_move"] == 30.0       # 10 iterations/ms x 3 ms
    assert games[0]["agent_move_stats"]["fast"]["avg_simulations_per_move"] == 40.0      # 20/ms x 2 ms

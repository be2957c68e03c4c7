"""Win-probability snapshot features on the device: bit-exact against the reference fixtures and
the CPU oracle, through the C ABI."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    from reinforcementlearning_blokus_b200 import device

    return device


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "snapshots.npz"))
    g = np.load(os.path.join(golden_dir, "games.npz"))
    return z, g


def test_replayed_games_match_reference_rows(dev, golden_dir):
    """The 16 golden games replayed on the device: capture points, boards, incremental frontier sets and all
    34 x 4 floats of every captured row equal what the unmodified reference produced."""
    z, g = _golden(golden_dir)
    ck = z["checkpoints"]
    r = dev.replay_snapshots(torch.from_numpy(g["traces"].copy()).cuda(), ck, int(z["max_turns"][0]), want_states=True)
    ti = r["turn_index"].cpu().numpy()
    assert (ti == z["turn_index"]).all()
    hit = ti >= 0
    assert (r["states"].cpu().numpy()[hit] == z["snap_states"][hit]).all()
    want_fr = np.unpackbits(z["frontiers"], axis=-1, bitorder="little")[..., :400]
    got_fr = np.unpackbits(r["frontiers"].cpu().numpy().view(np.uint8).reshape(*ti.shape, 4, 56), axis=-1, bitorder="little")[..., :400]
    assert (got_fr[hit] == want_fr[hit]).all()
    got = r["features"].cpu().numpy()
    assert (got[hit] == z["features"][hit]).all()
    assert (got[~hit] == 0).all()


def test_state_corpus_matches_reference_rows(dev, golden_dir):
    """The learned-evaluator call (rebuilt frontier sets, turn_index = move_count) on the 240-state corpus."""
    z, _ = _golden(golden_dir)
    states = np.load(os.path.join(golden_dir, "random_states.npz"))["states"]
    got = dev.snapshot_features(dev.to_device(states)).cpu().numpy()
    assert (got == z["state_features"]).all()


def test_explicit_frontiers_and_turn_index(dev, golden_dir):
    """Boards captured mid-game with their (stale-cell carrying) frontier sets and loop turn counts as inputs."""
    z, _ = _golden(golden_dir)
    hit = z["turn_index"] >= 0
    states = z["snap_states"][hit]
    fr_bits = np.unpackbits(z["frontiers"], axis=-1, bitorder="little")[hit]            # [M, 4, 400]
    fr_words = np.packbits(np.pad(fr_bits, ((0, 0), (0, 0), (0, 48))), axis=-1, bitorder="little").view("<i8").reshape(-1, 4, 7)
    got = dev.snapshot_features(dev.to_device(states), torch.from_numpy(fr_words.copy()).cuda(),
                                torch.from_numpy(z["turn_index"][hit].astype(np.int32)).cuda(),
                                int(z["max_turns"][0])).cpu().numpy()
    assert (got == z["features"][hit]).all()


def test_random_playouts_against_oracle(dev):
    """2,000 device playouts replayed on the device and on the CPU oracle: same capture points, same rows."""
    from oracle import orc

    n, ck = 2000, [0, 5, 8, 16, 24, 31, 32, 40, 48, 56, 64, 70]
    r = dev.playout(n, seed=2024, want_trace=True)
    rep = dev.replay_snapshots(r["trace"], ck, max_turns=300)
    traces = r["trace"].cpu().numpy()
    feats, ti = rep["features"].cpu().numpy(), rep["turn_index"].cpu().numpy()
    from concurrent.futures import ThreadPoolExecutor

    def check(gi):                                                # the oracle call releases the GIL
        o = orc.replay_snapshots(traces[gi], ck, max_turns=300)
        hit = ti[gi] >= 0
        return bool((o["turn_index"] == ti[gi]).all() and (o["features"][hit] == feats[gi][hit]).all())

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        ok = list(pool.map(check, range(0, n, 4)))                # 500 games = 6,000 boards on the CPU
    assert all(ok), [4 * i for i, v in enumerate(ok) if not v][:5]
    # size-independent properties over all games
    hit = ti >= 0
    col = {c: i for i, c in enumerate(__import__("reinforcementlearning_blokus_b200.winprob", fromlist=["x"]).SNAPSHOT_FEATURE_COLUMNS)}
    f = feats[hit]                                                 # [M, 4, 34]
    assert (f[..., col["phase_ply"]] == np.broadcast_to(np.array(ck)[None, :], ti.shape)[hit][:, None]).all()
    assert (f[..., col["pieces_used_count"]].sum(axis=1) == f[:, 0, col["phase_ply"]]).all()      # one piece per ply
    assert (f[..., col["pieces_used_count"]] + f[..., col["pieces_remaining_count"]] == 21).all()
    assert np.allclose(f[..., col["player_board_occupancy"]].sum(axis=1), f[:, 0, col["phase_board_occupancy"]])
    total = f[..., col["mobility_total_placements"]]
    assert (f[..., col["blocking_exposure_legal_moves_opp_sum"]] == total.sum(axis=1, keepdims=True) - total).all()
    assert (f[..., col["corner_differential"]] == 2 * f[..., col["frontier_size"]] - f[..., col["frontier_size"]].sum(axis=1, keepdims=True)).all()


def test_object_api_and_dataset_rows(dev, golden_dir, tmp_path):
    """winprob.extract_player_snapshot_features on a live BlokusGame (incremental frontier kept by the host Board) and
    the snapshots.csv rows of a device arena."""
    from reinforcementlearning_blokus_b200 import BlokusGame, Player, arena, winprob
    from reinforcementlearning_blokus_b200.move_generator import Move

    z, g = _golden(golden_dir)
    game = BlokusGame(enable_telemetry=False)
    turn = 0
    for t in range(int(g["turns"][0])):
        turn += 1
        if g["traces"][0, t] == 0xFFFF:
            game.board._update_current_player(); game._check_game_over()
            continue
        assert game.make_move(Move.from_id(int(g["traces"][0, t])), game.get_current_player())
        if game.board.move_count in (8, 32, 56):
            k = list(z["checkpoints"]).index(game.board.move_count)
            ctx = winprob.build_snapshot_runtime_context(game.board, turn_index=turn, max_turns=int(z["max_turns"][0]))
            for p in Player:
                f = winprob.coerce_feature_dict(winprob.extract_player_snapshot_features(game.board, player=p, context=ctx))
                assert [f[c] for c in winprob.SNAPSHOT_FEATURE_COLUMNS] == z["features"][0, k, p.value - 1].tolist()
    out = arena.run_device_arena({"a": "random", "b": "heuristic", "c": "random", "d": "heuristic"}, 24, run_seed=7,
                                 snapshots=winprob.DEFAULT_SNAPSHOT_PLYS)
    rows = out["snapshot_rows"]
    assert len(rows) == 4 * sum(len(r["snapshot_checkpoints_hit"]) for r in out["records"]) and len(rows) >= 24 * 4 * 6
    first = rows[0]
    assert list(first)[:7] == ["run_id", "game_id", "game_index", "game_seed", "checkpoint_index", "checkpoint_ply", "ply"]
    assert all(r["label_is_winner"] == int(r["player_id"] in out["records"][r["game_index"]]["winner_ids"]) for r in rows)
    meta = winprob.write_snapshots_dataset(str(tmp_path), rows)
    assert meta["enabled"] and meta["rows"] == len(rows) and os.path.exists(meta["path_csv"])
    assert winprob.LearnedWinProbabilityEvaluator("x/dummy_model.json").predict_player_win_probability(game.board, Player.RED) == 0.5


def test_learned_evaluator_with_sklearn_artifacts(dev, golden_dir, tmp_path):
    """mcts/learned_evaluator.py:97-152 with real joblib artifacts of both supported model types: the probability is
    the mean over the three opponents of predict_proba(feature_i - feature_j), features taken from the reference rows."""
    joblib = pytest.importorskip("joblib")
    from sklearn.ensemble import GradientBoostingClassifier
    from sklearn.linear_model import LogisticRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    from reinforcementlearning_blokus_b200 import Board, Player, winprob

    z, _ = _golden(golden_dir)
    states = np.load(os.path.join(golden_dir, "random_states.npz"))["states"].view(winprob.np.uint8).reshape(-1, 256)
    rows = z["state_features"]                                             # reference rows of those boards [240, 4, 34]
    rng = np.random.default_rng(0)
    x = np.concatenate([rows[:, 0] - rows[:, 1], rows[:, 2] - rows[:, 3]])
    y = rng.integers(0, 2, len(x))
    cols = winprob.SNAPSHOT_FEATURE_COLUMNS
    logreg = Pipeline([("scale", StandardScaler()), ("clf", LogisticRegression(max_iter=200))]).fit(x, y)
    gbt = GradientBoostingClassifier(n_estimators=5, random_state=0).fit(x, y)
    p1, p2 = str(tmp_path / "logreg.joblib"), str(tmp_path / "gbt.joblib")
    joblib.dump({"model_type": "pairwise_logreg", "feature_columns": cols, "pipeline": logreg}, p1)
    joblib.dump({"model_type": "pairwise_gbt_phase", "feature_columns": cols, "phase_models": {"mid": gbt}, "fallback_model": gbt}, p2)
    from reinforcementlearning_blokus_b200.state import STATE_DTYPE
    for path, model in ((p1, logreg), (p2, gbt)):
        ev = winprob.LearnedWinProbabilityEvaluator(path)
        for i in (3, 77, 200):
            board = Board.from_packed(states[i].view(STATE_DTYPE)[0])
            for pl in (Player.RED, Player.GREEN):
                a = pl.value - 1
                want = float(np.mean([model.predict_proba((rows[i, a] - rows[i, b])[None])[:, 1][0] for b in range(4) if b != a]))
                assert ev.predict_player_win_probability(board, pl) == pytest.approx(want, abs=1e-15)
        assert 0.0 < ev.potential(board, Player.BLUE) < 1.0
    with pytest.raises(ValueError):
        joblib.dump({"model_type": "other"}, str(tmp_path / "bad.joblib"))
        winprob.LearnedWinProbabilityEvaluator(str(tmp_path / "bad.joblib"))


def test_replay_edge_cases(dev):
    """Illegal or out-of-range ids stop the replay of that game only; replay from given start states; empty batch."""
    from oracle import orc

    r = dev.playout(64, seed=5, want_trace=True)
    tr = r["trace"].cpu().numpy().copy()
    ck = [0, 8, 16, 24, 32]
    first = np.array([np.nonzero(t != 0xFFFF)[0][20] for t in tr])          # turn of the 21st placement
    tr[0, first[0]] = 40000                                                  # out of range
    tr[1, first[1]] = tr[1, np.nonzero(tr[1] != 0xFFFF)[0][0]]               # the game's first move again: illegal now
    rep = dev.replay_snapshots(torch.from_numpy(tr).cuda(), ck, max_turns=100)
    ti = rep["turn_index"].cpu().numpy()
    assert (ti[:2, :3] >= 0).all() and (ti[:2, 3:] == -1).all()              # plies 0, 8, 16 captured, 24 and 32 not
    assert (ti[2:] >= 0).all()
    for gi in (0, 1, 2):
        o = orc.replay_snapshots(tr[gi], ck, max_turns=100)
        assert (o["turn_index"] == ti[gi]).all()
        assert (o["features"][ti[gi] >= 0] == rep["features"][gi].cpu().numpy()[ti[gi] >= 0]).all()
    # continue games from mid-game boards: turns and plies count from the given start state
    seg = dev.playout(32, seed=9, max_turns=20, want_final=True, want_trace=True)
    rest = dev.playout(32, seed=9, start_states=seg["final_states"], first_turn=20, want_trace=True)
    tail = rest["trace"][:, 20:].contiguous()
    mc = seg["final_states"].cpu().numpy().view(orc.STATE_DTYPE).reshape(-1)
    ck2 = sorted({int(mc["move_count"].min()), int(mc["move_count"].max()), 30, 40})
    rep = dev.replay_snapshots(tail, ck2, max_turns=76, start=seg["final_states"])
    ti = rep["turn_index"].cpu().numpy()
    for gi in range(0, 32, 5):
        o = orc.replay_snapshots(tail[gi].cpu().numpy(), ck2, max_turns=76, start=mc[gi])
        assert (o["turn_index"] == ti[gi]).all()
        hit = ti[gi] >= 0
        assert hit.any() and (o["features"][hit] == rep["features"][gi].cpu().numpy()[hit]).all()
    empty = dev.snapshot_features(torch.empty((0, 256), dtype=torch.uint8, device="cuda"))
    assert tuple(empty.shape) == (0, 4, 34)
    with pytest.raises(ValueError):
        dev.replay_snapshots(tail, [])


def test_pairwise_tensors_equal_dataframe_path(dev):
    """The training set straight from device tensors equals the data-frame route of analytics/winprob/dataset.py."""
    import pandas as pd

    from reinforcementlearning_blokus_b200 import arena, winprob

    out = arena.run_device_arena({"a": "random", "b": "heuristic", "c": "random", "d": "random"}, 40, run_seed=3,
                                 snapshots=[8, 24, 48, 64])
    frame, meta = winprob.build_pairwise_dataset(pd.DataFrame(out["snapshot_rows"]))
    scores = torch.tensor([[r["final_scores"][str(p)] for p in (1, 2, 3, 4)] for r in out["records"]])
    x, y, g, c = winprob.pairwise_tensors(out["snapshot_features"], scores)
    assert x.shape[0] == meta["rows"] > 0
    assert np.array_equal(x.cpu().numpy(), frame[winprob.SNAPSHOT_FEATURE_COLUMNS].to_numpy(dtype=np.float64))
    assert np.array_equal(y.cpu().numpy(), frame["label"].to_numpy())
    assert np.array_equal(c.cpu().numpy(), frame["checkpoint_index"].to_numpy())
    assert [f"device_arena_g{i:04d}" for i in g.cpu().tolist()] == frame["game_id"].tolist()

"""K5 parity: the device UCT kernel against the sequential C oracle (same algorithm restated from
mcts/mcts_agent.py, same canonical untried order, same Philox rollout stream)."""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402
from reinforcementlearning_blokus_b200 import Board, MCTSAgent, Move, Player, RandomAgent, ZobristHash, get_shared_generator  # noqa: E402
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.mcts import DeviceUCT, TranspositionTable  # noqa: E402
from reinforcementlearning_blokus_b200.state import STATE_DTYPE, empty_states  # noqa: E402


def _roots(golden_dir):
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    states = z["states"].view(STATE_DTYPE).reshape(-1)
    picks = [s for s in states if 4 <= int(s["move_count"]) <= 40][:3]
    return [empty_states(1)[0]] + picks


@pytest.mark.parametrize("variant", ["uct", "flat"])
def test_search_matches_oracle(golden_dir, variant):
    kw = dict(max_depth=0, ucb_variant=0, reward_mode=0) if variant == "uct" else dict(max_depth=1, ucb_variant=1, reward_mode=1)
    iters = 160
    for ri, root in enumerate(_roots(golden_dir)):
        for tree in (0, 5):
            want = orc.mcts_search(root, iters, seed=99, tree_id=tree, **kw)
            eng = DeviceUCT(iters, n_trees=1, seed=99, tree_offset=tree, **kw)
            res = eng.search(dev.to_device(np.atleast_1d(root)), iters)
            n = int(res["counts"][0])
            assert n == want["n_root"]
            assert int(res["status"][0]) == 0
            assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n]), (variant, ri, tree)
            assert np.array_equal(res["reward"][0, :n].cpu().numpy(), want["reward"][:n])
            assert np.array_equal(res["moves"][0, :n].cpu().numpy(), want["moves"][:n])
            assert int(res["visits"][0, :n].sum()) == iters


def test_resume_equals_one_shot_and_trees_are_independent(golden_dir):
    root = dev.to_device(np.atleast_1d(_roots(golden_dir)[1]))
    one = DeviceUCT(192, n_trees=4, seed=5).search(root, 192)
    v1 = one["visits"].clone(); w1 = one["reward"].clone()
    eng = DeviceUCT(192, n_trees=4, seed=5)
    eng.search(root, 64)
    eng.search(None, 64, resume=True)
    res = eng.search(None, 64, resume=True)
    assert torch.equal(res["visits"], v1) and torch.equal(res["reward"], w1)
    # tree k of a 4-tree batch == a single tree with tree_offset k
    solo = DeviceUCT(192, n_trees=1, seed=5, tree_offset=2).search(root, 192)
    assert torch.equal(solo["visits"][0], v1[2]) and torch.equal(solo["reward"][0], w1[2])
    assert not torch.equal(w1[0], w1[1])        # different Philox streams -> different rollouts


def test_transposition_table_matches_oracle(golden_dir):
    """Rewards cached by Zobrist key are reused by a second search of the same agent (mcts_agent.py:418-434)."""
    root_np = _roots(golden_dir)[2]
    zob = ZobristHash(seed=42)
    assert np.array_equal(zob.table, orc.zobrist_table(42))
    tt = TranspositionTable(max_size=1 << 16)
    ott = orc.OracleTT(18)
    root = dev.to_device(np.atleast_1d(root_np))
    for rep in range(2):                      # second pass hits the entries of the first
        want = orc.mcts_search(root_np, 120, seed=8, tree_id=0, tt=ott, zobrist=zob.table)
        res = DeviceUCT(120, n_trees=1, seed=8, tt=tt.device, zobrist=zob).search(root, 120)
        n = want["n_root"]
        assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n]), rep
        assert np.array_equal(res["reward"][0, :n].cpu().numpy(), want["reward"][:n]), rep
        assert len(tt) == ott.size()
    ott.close()


def test_mcts_agent_interface(golden_dir):
    board = Board.from_packed(_roots(golden_dir)[1])
    player = board.current_player
    gen = get_shared_generator()
    legal = gen.get_legal_moves(board, player)
    agent = MCTSAgent(iterations=150, seed=4)
    mv = agent.select_action(board, player, legal)
    assert mv in legal
    assert agent.stats["iterations_run"] == 150
    info = agent.get_action_info()
    assert info["type"] == "mcts" and "transposition_stats" in info
    # the move equals the oracle's best-by-visits with the reference's tie-break (first max in creation order)
    st = board.to_packed()
    ott = orc.OracleTT(18)
    want = orc.mcts_search(st, 150, seed=4, tree_id=0, rollout_policy=1, tt=ott, zobrist=orc.zobrist_table(4))
    n = want["n_root"]
    best, best_v = -1, -1
    for rank in range(n - 1, -1, -1):
        if want["visits"][rank] > best_v:
            best, best_v = rank, want["visits"][rank]
    assert mv.to_id() == sorted(m.to_id() for m in legal)[best]
    assert agent.select_action(board, player, []) is None
    assert agent.select_action(board, player, legal[:1]) == legal[0]
    agent.reset()
    with pytest.raises(ValueError):
        MCTSAgent(max_rollout_moves=0)
    # time-limited search runs and returns a legal move
    timed = MCTSAgent(time_limit=0.2, seed=1, use_transposition_table=False)
    assert timed.select_action(board, player, legal) in legal
    assert timed.stats["iterations_run"] >= 64
    # root-parallel trees: visits add up over trees
    par = MCTSAgent(iterations=64, seed=2, num_trees=8, use_transposition_table=False)
    r = par.search_root(board, player)
    assert int(r["visits"].sum()) == 64 * 8


def test_playout_segments_chain(golden_dir):
    """A game played in segments through final_states equals the one-shot game (bk_playout first_turn)."""
    n = 128
    full = dev.playout(n, seed=9, game_offset=3, want_trace=True, want_final=True)
    st, tr = None, None
    cuts = [0, 7, 20, 45, 96]
    out = {}
    for a, b in zip(cuts[:-1], cuts[1:]):
        r = dev.playout(n, seed=9, game_offset=3, start_states=st, first_turn=a, max_turns=b, want_trace=True,
                        want_final=True, out=out)
        st = r["final_states"].clone()
    assert torch.equal(r["trace"], full["trace"])
    assert torch.equal(r["game_scores"], full["game_scores"])
    assert torch.equal(st, full["final_states"])


def test_heuristic_policy_playouts_and_rollouts(golden_dir):
    """N1: HeuristicAgent softmax as a device policy -- whole games and UCT rollouts equal the oracle's."""
    n, seed = 96, 31
    r = dev.playout(n, seed=seed, game_offset=500, policy=dev.POLICY_HEURISTIC, want_trace=True, want_choice=True,
                    want_final=True)
    tr, ch, gs = r["trace"].cpu().numpy(), r["choice"].cpu().numpy(), r["game_scores"].cpu().numpy()
    for i in range(0, n, 6):
        o = orc.playout(seed, 500 + i, policy=1, want_final=True)
        assert np.array_equal(tr[i], o["trace"]), i
        assert np.array_equal(ch[i], o["choice"]) and np.array_equal(gs[i], o["game_scores"])
        assert r["final_states"][i].cpu().numpy().tobytes() == o["final"].tobytes()
    # the heuristic prefers big pieces: it scores clearly higher than uniform play
    uni = dev.playout(n, seed=seed, game_offset=500)["game_scores"].float().mean().item()
    assert gs.mean() > uni + 3
    # UCT with heuristic rollouts (the reference's default rollout policy)
    root = _roots(golden_dir)[1]
    want = orc.mcts_search(root, 80, seed=12, tree_id=3, rollout_policy=1)
    res = DeviceUCT(80, n_trees=1, seed=12, tree_offset=3, rollout_policy=1).search(dev.to_device(np.atleast_1d(root)), 80)
    nr = want["n_root"]
    assert np.array_equal(res["visits"][0, :nr].cpu().numpy(), want["visits"][:nr])
    assert np.array_equal(res["reward"][0, :nr].cpu().numpy(), want["reward"][:nr])


def test_learned_evaluator_options_with_dummy_artifact(golden_dir):
    """mcts_agent.py:427-466 with the reference's dummy artifact (every probability 0.5): leaf evaluation replaces the
    rollout by a constant, progressive bias and (gamma = 1) shaping terms vanish.  Kernel == oracle, with and without
    the transposition table; the agent accepts the reference's constructor arguments."""
    iters = 200
    for root_np in _roots(golden_dir)[:2]:
        root = dev.to_device(np.atleast_1d(root_np))
        want = orc.mcts_search(root_np, iters, seed=11, tree_id=3, reward_mode=2)
        res = DeviceUCT(iters, n_trees=1, seed=11, tree_offset=3, reward_mode=2).search(root, iters)
        n = want["n_root"]
        assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n])
        assert np.array_equal(res["reward"][0, :n].cpu().numpy(), want["reward"][:n])
        assert np.array_equal(res["reward"][0, :n].cpu().numpy(), want["visits"][:n])          # one half per visit
        zob, tt, ott = ZobristHash(seed=42), TranspositionTable(max_size=1 << 16), orc.OracleTT(18)
        want = orc.mcts_search(root_np, iters, seed=11, tree_id=3, reward_mode=2, tt=ott, zobrist=zob.table)
        res = DeviceUCT(iters, n_trees=1, seed=11, tree_offset=3, reward_mode=2, tt=tt.device, zobrist=zob).search(root, iters)
        assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n])
        ott.close()
    board = Board.from_packed(_roots(golden_dir)[1])
    player = board.current_player
    legal = get_shared_generator().get_legal_moves(board, player)
    agent = MCTSAgent(iterations=120, seed=4, learned_model_path="models/dummy_model.json", leaf_evaluation_enabled=True,
                      progressive_bias_enabled=True, potential_shaping_enabled=True, use_transposition_table=False)
    r = agent.search_root(board, player)
    assert int(r["visits"].sum()) == 120 and np.array_equal(r["reward"], r["visits"] * 0.5)
    assert agent.stats["leaf_eval_calls"] == 120 and agent.stats["progressive_bias_updates"] == 120
    assert agent.select_action(board, player, legal) in legal
    # bias / shaping only: rollouts as usual, same tree as the plain agent
    plain = MCTSAgent(iterations=120, seed=4, use_transposition_table=False).search_root(board, player)
    shaped = MCTSAgent(iterations=120, seed=4, use_transposition_table=False, learned_model_path="x/dummy_model.json",
                       progressive_bias_enabled=True, potential_shaping_enabled=True, potential_mode="logit",
                       potential_shaping_gamma=0.9).search_root(board, player)
    assert np.array_equal(plain["visits"], shaped["visits"]) and np.array_equal(plain["reward"], shaped["reward"])
    with pytest.raises(ValueError):
        MCTSAgent(leaf_evaluation_enabled=True)
    with pytest.raises(NotImplementedError):
        MCTSAgent(learned_model_path="x/dummy_model.json", potential_shaping_enabled=True, potential_shaping_gamma=0.9)


def test_leaf_batched_rollouts_match_oracle(golden_dir):
    """BASELINE config 4 "GPU-batched leaf rollouts": a CTA of W warps owns the tree, every iteration rolls out W times
    from the new leaf (streams 1..W of the tree) and backs the rewards up together.  Kernel == sequential oracle for
    W = 2, 4, 8, 16 (UCT, flat bandit, heuristic rollouts, constant leaf); W = 1 is the unbatched search; resume works."""
    iters = 96
    roots = _roots(golden_dir)
    for W, kw in ((2, dict()), (4, dict(max_depth=1, ucb_variant=1, reward_mode=1)), (8, dict()), (16, dict()),
                  (4, dict(rollout_policy=1)), (16, dict(rollout_policy=1)), (8, dict(reward_mode=2))):
        for ri in (0, 2):
            root = dev.to_device(np.atleast_1d(roots[ri]))
            want = orc.mcts_search(roots[ri], iters, seed=21, tree_id=ri, leaf_batch=W, **kw)
            res = DeviceUCT(iters, n_trees=1, seed=21, tree_offset=ri, leaf_batch=W, **kw).search(root, iters)
            n = want["n_root"]
            assert int(res["counts"][0]) == n and int(res["status"][0]) == 0
            assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n]), (W, kw, ri)
            assert np.array_equal(res["reward"][0, :n].cpu().numpy(), want["reward"][:n]), (W, kw, ri)
            assert np.array_equal(res["moves"][0, :n].cpu().numpy(), want["moves"][:n])
            assert int(res["visits"][0, :n].sum()) == iters * W
    root = dev.to_device(np.atleast_1d(roots[1]))
    plain = DeviceUCT(iters, n_trees=3, seed=5).search(root, iters)
    one = DeviceUCT(iters, n_trees=3, seed=5, leaf_batch=1).search(root, iters)
    assert torch.equal(plain["visits"], one["visits"]) and torch.equal(plain["reward"], one["reward"])
    full = DeviceUCT(iters, n_trees=5, seed=6, leaf_batch=4).search(root, iters)
    v, w = full["visits"].clone(), full["reward"].clone()
    eng = DeviceUCT(iters, n_trees=5, seed=6, leaf_batch=4)
    eng.search(root, 32); eng.search(None, 40, resume=True)
    res = eng.search(None, 24, resume=True)
    assert torch.equal(res["visits"], v) and torch.equal(res["reward"], w)
    # with a transposition table a batch caches the sum of its rewards; a second search hits the entries of the first
    zob, tt, ott = ZobristHash(seed=42), TranspositionTable(max_size=1 << 16), orc.OracleTT(18)
    root_np = roots[2]
    for rep in range(2):
        want = orc.mcts_search(root_np, 80, seed=8, tree_id=0, leaf_batch=4, tt=ott, zobrist=zob.table)
        res = DeviceUCT(80, n_trees=1, seed=8, leaf_batch=4, tt=tt.device, zobrist=zob).search(dev.to_device(np.atleast_1d(root_np)), 80)
        n = want["n_root"]
        assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n]), rep
        assert np.array_equal(res["reward"][0, :n].cpu().numpy(), want["reward"][:n]), rep
        assert len(tt) == ott.size()
    ott.close()
    # the agent: same budget in simulations, spent 8 at a time
    from reinforcementlearning_blokus_b200 import FastMCTSAgent
    board = Board.from_packed(roots[1])
    legal = get_shared_generator().get_legal_moves(board, board.current_player)
    out = FastMCTSAgent(iterations=800, time_limit=5.0, seed=3, leaf_batch=8).think(board, board.current_player, legal, 5000)
    assert out["move"] in legal and out["stats"]["nodesEvaluated"] == 800
    assert sum(m["visits"] for m in out["stats"]["topMoves"]) <= 800


def test_leaf_batched_many_trees_per_cta(golden_dir):
    """More leaf-batched trees than resident CTAs: every CTA takes several trees from the ticket counter, and tree k of
    the batch still equals a lone tree with tree_offset k."""
    root = dev.to_device(np.atleast_1d(_roots(golden_dir)[2]))
    n_trees, iters, W = 400, 12, 16
    res = DeviceUCT(iters, n_trees=n_trees, seed=31, leaf_batch=W).search(root, iters)
    assert int(res["status"].max()) == 0
    assert (res["visits"].sum(dim=1) == iters * W).all()
    for k in (0, 147, 148, 399):
        solo = DeviceUCT(iters, n_trees=1, seed=31, tree_offset=k, leaf_batch=W).search(root, iters)
        assert torch.equal(solo["visits"][0], res["visits"][k]) and torch.equal(solo["reward"][0], res["reward"][k])


def test_batched_mcts_selfplay():
    """batched.selfplay_mcts: every move of every game chosen by device UCT.  Games are legal and complete (the oracle
    replays the traces to the same scores), deterministic in the seed, and the first move of each game is the oracle
    search's most visited move under the reference's tie rule."""
    from reinforcementlearning_blokus_b200 import batched

    n, iters = 12, 48
    a = batched.selfplay_mcts(n, iters, seed=13)
    b = batched.selfplay_mcts(n, iters, seed=13)
    assert torch.equal(a["trace"], b["trace"]) and torch.equal(a["game_scores"], b["game_scores"])
    tr, gs, turns = a["trace"].cpu().numpy(), a["game_scores"].cpu().numpy(), a["turns"].cpu().numpy()
    assert int(dev.has_moves(a["final_states"]).max()) == 0                       # all games finished
    for g in range(n):
        board = orc.OracleBoard()
        for t in range(int(turns[g])):
            if tr[g, t] == 0xFFFF:
                assert not board.has_legal_moves(board.current_player)            # passes only when forced
                board.pass_turn()
            else:
                assert board.make_move(board.current_player, int(tr[g, t]))
        assert [board.game_score(p) for p in range(4)] == gs[g].tolist()
        want = orc.mcts_search(empty_states(1)[0], iters, seed=13, tree_id=g)
        m = want["n_root"]
        best = max(range(m - 1, -1, -1), key=lambda r: want["visits"][r])          # highest rank among the most visited
        assert int(tr[g, 0]) == int(want["moves"][best])
    # root-parallel trees per game and leaf-batched rollouts run too, and return visit-count policy targets
    c = batched.selfplay_mcts(4, 16, seed=2, trees_per_game=3, leaf_batch=4, want_policy=True, max_turns=6)
    v0, m0 = c["policy"][0]
    assert v0.shape[0] == 4 and int(v0[0].sum()) == 3 * 16 * 4 and len(c["policy"]) == 6
    # a game's streams depend on its GLOBAL index only: games 2..3 of a 4-game batch == a 2-game batch at offset 2
    d = batched.selfplay_mcts(2, 16, seed=2, trees_per_game=3, leaf_batch=4, max_turns=6, game_offset=2)
    assert torch.equal(c["trace"][2:, :6], d["trace"][:, :6])
    # a finished game is left alone while the slower games of the batch go on (its side to move does not keep rotating)
    e = batched.selfplay_mcts(6, 8, seed=21)
    solo = [batched.selfplay_mcts(1, 8, seed=21, game_offset=g) for g in range(6)]
    for g in range(6):
        assert torch.equal(e["final_states"][g], solo[g]["final_states"][0]) and torch.equal(e["trace"][g], solo[g]["trace"][0])


def _logreg_artifact(golden_dir, tmp_path, scaler=True):
    joblib = pytest.importorskip("joblib")
    from sklearn.linear_model import LogisticRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    from reinforcementlearning_blokus_b200 import winprob

    rows = np.load(os.path.join(golden_dir, "snapshots.npz"))["state_features"]            # reference rows [240, 4, 34]
    x = np.concatenate([rows[:, 0] - rows[:, 1], rows[:, 2] - rows[:, 3], rows[:, 1] - rows[:, 2]])
    # a learnable target: the side with more mobility and more squares placed tends to win
    c = winprob.SNAPSHOT_FEATURE_COLUMNS
    y = (x[:, c.index("mobility_total_placements")] + 20 * x[:, c.index("player_board_occupancy")] * 400 > 0).astype(int)
    steps = ([("scale", StandardScaler())] if scaler else []) + [("clf", LogisticRegression(max_iter=500))]
    pipe = Pipeline(steps).fit(x, y)
    path = str(tmp_path / ("logreg_scaled.joblib" if scaler else "logreg.joblib"))
    joblib.dump({"model_type": "pairwise_logreg", "feature_columns": c, "pipeline": pipe}, path)
    return path, pipe, rows


def test_learned_leaf_evaluation_in_the_tree(golden_dir, tmp_path):
    """mcts_agent.py:427-451 with a fitted pairwise logistic model: the kernel computes the leaf's snapshot features and
    the folded model itself.  (1) the folded weights reproduce scikit-learn's predict_proba; (2) kernel == sequential
    oracle (visits, moves, double reward sums) with and without root frontier sets and the transposition table;
    (3) the agent accepts the reference's constructor arguments and its leaf values are the evaluator's probabilities."""
    from reinforcementlearning_blokus_b200 import winprob

    path, pipe, rows = _logreg_artifact(golden_dir, tmp_path)
    ev = winprob.LearnedWinProbabilityEvaluator(path)
    w = ev.device_model
    assert w is not None and w.shape == (35,)
    d = rows[7, 0] - rows[7, 2]
    assert 1.0 / (1.0 + np.exp(-(w[34] + float(np.dot(w[:34], d))))) == pytest.approx(float(pipe.predict_proba(d[None])[0, 1]), abs=1e-12)
    iters = 150
    z = np.load(os.path.join(golden_dir, "snapshots.npz"))
    for ri, root_np in enumerate(_roots(golden_dir)[:3]):
        root = dev.to_device(np.atleast_1d(root_np))
        want = orc.mcts_search_learned(root_np, iters, seed=17, model=w, tree_id=ri)
        res = DeviceUCT(iters, n_trees=1, seed=17, tree_offset=ri, reward_mode=3, leaf_model=torch.from_numpy(w)).search(root, iters)
        n = want["n_root"]
        assert int(res["counts"][0]) == n and int(res["status"][0]) == 0
        assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n]), ri
        assert np.array_equal(res["moves"][0, :n].cpu().numpy(), want["moves"][:n])
        assert np.allclose(res["reward"][0, :n].cpu().numpy(), want["reward"][:n], rtol=0, atol=1e-12)
        assert int(res["visits"][0, :n].sum()) == iters
    # a mid-game board captured by the reference WITH its incremental frontier sets (stale cells included)
    g, k = 3, 4
    st = z["snap_states"][g, k].view(STATE_DTYPE)[0]
    fr_bits = np.unpackbits(z["frontiers"][g, k], axis=-1, bitorder="little")[:, :400]
    fr_words = np.packbits(np.pad(fr_bits, ((0, 0), (0, 48))), axis=-1, bitorder="little").view("<i8").reshape(1, 4, 7)
    want = orc.mcts_search_learned(st, iters, seed=4, model=w, root_frontier=fr_bits)
    eng = DeviceUCT(iters, n_trees=1, seed=4, reward_mode=3, leaf_model=torch.from_numpy(w))
    res = eng.search(dev.to_device(np.atleast_1d(st)), iters, root_frontiers=torch.from_numpy(fr_words.copy()))
    n = want["n_root"]
    assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n])
    assert np.allclose(res["reward"][0, :n].cpu().numpy(), want["reward"][:n], rtol=0, atol=1e-12)
    # transposition table: cached probabilities are reused by a second search
    zob, tt, ott = ZobristHash(seed=42), TranspositionTable(max_size=1 << 16), orc.OracleTT(18)
    for rep in range(2):
        want = orc.mcts_search_learned(st, 100, seed=9, model=w, tt=ott, zobrist=zob.table)
        res = DeviceUCT(100, n_trees=1, seed=9, reward_mode=3, leaf_model=torch.from_numpy(w), tt=tt.device, zobrist=zob).search(
            dev.to_device(np.atleast_1d(st)), 100)
        assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n]), rep
    ott.close()
    # the agent: leaf values are the evaluator's own probabilities of the children, for the player to move there
    board = Board.from_packed(_roots(golden_dir)[1])
    player = board.current_player
    legal = get_shared_generator().get_legal_moves(board, player)
    agent = MCTSAgent(iterations=len(legal), seed=2, learned_model_path=path, leaf_evaluation_enabled=True,
                      use_transposition_table=False)
    r = agent.search_root(board, player)
    assert int(r["visits"].sum()) == len(legal) and (r["visits"] == 1).all()          # one visit per root move
    ids = sorted(m.to_id() for m in legal)
    for rank in (0, len(ids) // 2, len(ids) - 1):
        child = board.copy()
        mv = Move.from_id(ids[rank])
        cells = get_shared_generator().piece_position_cache[mv.piece_id][mv.orientation]
        from reinforcementlearning_blokus_b200 import Position
        child.place_piece([Position(mv.anchor_row + a, mv.anchor_col + b) for a, b in cells], player, mv.piece_id, validate=False)
        assert r["reward"][rank] == pytest.approx(ev.predict_player_win_probability(child, child.current_player), abs=1e-9)
    assert agent.select_action(board, player, legal) in legal and agent.stats["leaf_eval_calls"] == len(legal)
    with pytest.raises(NotImplementedError):
        MCTSAgent(learned_model_path=path, progressive_bias_enabled=True)            # needs leaf evaluation too
    # potential-based shaping (mcts_agent.py:486-551): rollouts stay, truncated ones get w * (gamma * Phi(end) - Phi(leaf))
    for pol, mode, gamma in ((0, 0, 1.0), (0, 1, 0.9), (1, 0, 0.95)):
        root_np = _roots(golden_dir)[2]
        root = dev.to_device(np.atleast_1d(root_np))
        kw = dict(max_rollout_moves=6, rollout_policy=pol, shaping_weight=1.5, shaping_gamma=gamma)
        want = orc.mcts_search_shaped(root_np, 120, seed=31, model=w, potential_mode=mode, **kw)
        res = DeviceUCT(120, n_trees=1, seed=31, reward_mode=4, leaf_model=torch.from_numpy(w),
                        potential_mode="logit" if mode else "prob", **kw).search(root, 120)
        n = want["n_root"]
        assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n]), (pol, mode)
        assert np.allclose(res["reward"][0, :n].cpu().numpy(), want["reward"][:n], rtol=0, atol=1e-9)
        frac = res["reward"][0, :n].cpu().numpy()
        assert np.abs(frac - np.round(frac)).max() > 1e-6                      # the shaping term is really in the rewards
    shaped = MCTSAgent(iterations=60, seed=2, learned_model_path=path, potential_shaping_enabled=True,
                       potential_shaping_gamma=0.9, progressive_bias_enabled=True, use_transposition_table=False,
                       max_rollout_moves=8)
    assert shaped.select_action(board, player, legal) in legal
    # progressive bias (mcts_agent.py:88, 453-468) on top of leaf evaluation: kernel == oracle, and it changes the search
    for ri, root_np in enumerate(_roots(golden_dir)[1:3]):
        root = dev.to_device(np.atleast_1d(root_np))
        want = orc.mcts_search_learned(root_np, 300, seed=23, model=w, tree_id=ri, bias_weight=0.25)
        res = DeviceUCT(300, n_trees=1, seed=23, tree_offset=ri, reward_mode=3, leaf_model=torch.from_numpy(w),
                        progressive_bias_weight=0.25).search(root, 300)
        n = want["n_root"]
        assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n]), ri
        assert np.allclose(res["reward"][0, :n].cpu().numpy(), want["reward"][:n], rtol=0, atol=1e-12)
    late = z["snap_states"][0, 7].view(STATE_DTYPE)[0]                       # ply 56: few root moves, so selection happens
    plain = orc.mcts_search_learned(late, 300, seed=23, model=w)
    strong = orc.mcts_search_learned(late, 300, seed=23, model=w, bias_weight=5.0)
    assert 0 < plain["n_root"] < 150 and not np.array_equal(plain["visits"], strong["visits"])
    res = DeviceUCT(300, n_trees=1, seed=23, reward_mode=3, leaf_model=torch.from_numpy(w), progressive_bias_weight=5.0).search(
        dev.to_device(np.atleast_1d(late)), 300)
    assert np.array_equal(res["visits"][0, :plain["n_root"]].cpu().numpy(), strong["visits"][:plain["n_root"]])
    biased = MCTSAgent(iterations=120, seed=2, learned_model_path=path, leaf_evaluation_enabled=True,
                       progressive_bias_enabled=True, progressive_bias_weight=0.25, use_transposition_table=False)
    assert biased.select_action(board, player, legal) in legal and biased.stats["progressive_bias_updates"] == 120


def test_learned_gbt_phase_models_in_the_tree(golden_dir, tmp_path):
    """The other artifact type of the reference's evaluator, "pairwise_gbt_phase" (mcts/learned_evaluator.py:112-127):
    flattened scikit-learn gradient-boosted trees are walked inside the kernel.  Kernel == oracle for leaf evaluation
    (+ progressive bias) and for rollouts with potential shaping; the agent's leaf values equal scikit-learn's."""
    joblib = pytest.importorskip("joblib")
    from sklearn.ensemble import GradientBoostingClassifier

    from reinforcementlearning_blokus_b200 import Position, winprob

    rows = np.load(os.path.join(golden_dir, "snapshots.npz"))["state_features"]
    cols = winprob.SNAPSHOT_FEATURE_COLUMNS
    x = np.concatenate([rows[:, 0] - rows[:, 1], rows[:, 2] - rows[:, 3], rows[:, 1] - rows[:, 2]])
    y = (x[:, cols.index("mobility_total_placements")] + 8000 * x[:, cols.index("player_board_occupancy")] > 0).astype(int)
    occ = np.concatenate([(rows[:, a, 29] + rows[:, b, 29]) / 2 for a, b in ((0, 1), (2, 3), (1, 2))])
    early = GradientBoostingClassifier(n_estimators=45, max_depth=3, random_state=1).fit(x[occ < 0.4], y[occ < 0.4])
    rest = GradientBoostingClassifier(n_estimators=30, max_depth=2, random_state=2).fit(x, y)
    path = str(tmp_path / "gbt.joblib")
    joblib.dump({"model_type": "pairwise_gbt_phase", "feature_columns": cols, "phase_models": {"early": early},
                 "fallback_model": rest}, path)
    ev = winprob.LearnedWinProbabilityEvaluator(path)
    blob = ev.device_model
    assert ev.device_model_kind == 1 and blob is not None
    tb = torch.from_numpy(blob)
    for ri, root_np in enumerate(_roots(golden_dir)[:3]):
        root = dev.to_device(np.atleast_1d(root_np))
        want = orc.mcts_search_learned(root_np, 120, seed=41, model=blob, model_kind=1, tree_id=ri, bias_weight=0.25)
        res = DeviceUCT(120, n_trees=1, seed=41, tree_offset=ri, reward_mode=3, leaf_model=tb, leaf_model_kind=1,
                        progressive_bias_weight=0.25).search(root, 120)
        n = want["n_root"]
        assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n]), ri
        assert np.allclose(res["reward"][0, :n].cpu().numpy(), want["reward"][:n], rtol=0, atol=1e-12)
    root_np = _roots(golden_dir)[2]
    kw = dict(max_rollout_moves=5, rollout_policy=1, shaping_weight=2.0, shaping_gamma=0.9)
    want = orc.mcts_search_shaped(root_np, 80, seed=5, model=blob, model_kind=1, potential_mode=1, **kw)
    res = DeviceUCT(80, n_trees=1, seed=5, reward_mode=4, leaf_model=tb, leaf_model_kind=1, potential_mode="logit", **kw).search(
        dev.to_device(np.atleast_1d(root_np)), 80)
    n = want["n_root"]
    assert np.array_equal(res["visits"][0, :n].cpu().numpy(), want["visits"][:n])
    assert np.allclose(res["reward"][0, :n].cpu().numpy(), want["reward"][:n], rtol=0, atol=1e-9)
    board = Board.from_packed(_roots(golden_dir)[1])
    player = board.current_player
    legal = get_shared_generator().get_legal_moves(board, player)
    agent = MCTSAgent(iterations=len(legal), seed=2, learned_model_path=path, leaf_evaluation_enabled=True,
                      use_transposition_table=False)
    r = agent.search_root(board, player)
    ids = sorted(m.to_id() for m in legal)
    for rank in (0, len(ids) // 3, len(ids) - 1):
        child = board.copy()
        mv = Move.from_id(ids[rank])
        cells = get_shared_generator().piece_position_cache[mv.piece_id][mv.orientation]
        child.place_piece([Position(mv.anchor_row + a, mv.anchor_col + b) for a, b in cells], player, mv.piece_id, validate=False)
        assert r["reward"][rank] == pytest.approx(ev.predict_player_win_probability(child, child.current_player), abs=1e-9)


def test_config4_fast_mcts_1000_sims_on_kat1_roots(golden_dir):
    """BASELINE config 4 as SURVEY 8d C4 spells it out: roots = the five KAT-1 benchmark states and the empty board,
    FastMCTS-interface agent with 1,000 simulations per move (leaf-batched rollouts), device transposition keys from the
    reference-order Zobrist table.  Checks: move in the legal set, stats keys of fast_mcts_agent.py:204-225, root visits
    add up to the simulations, seed determinism, TT keys == ZobristHash(seed).hash_board on the roots."""
    import json

    from reinforcementlearning_blokus_b200 import FastMCTSAgent

    with open(os.path.join(golden_dir, "kat1_states.json")) as f:
        kat = json.load(f)
    roots = [empty_states(1)[0]] + [np.frombuffer(bytes.fromhex(k["packed_hex"]), dtype=STATE_DTYPE)[0] for k in kat]
    zob = ZobristHash(seed=42)
    keys = zob.hash_states(dev.to_device(np.array(roots))).cpu().numpy().view(np.uint64)
    gen = get_shared_generator()
    for i, st in enumerate(roots):
        board = Board.from_packed(st)
        player = board.current_player
        assert zob.hash_board(board) == int(keys[i])
        legal = gen.get_legal_moves(board, player)
        outs = [FastMCTSAgent(iterations=1000, time_limit=30.0, seed=7, leaf_batch=8).think(board, player, legal, 30000)
                for _ in range(2)]
        a, b = outs
        assert a["move"] in legal and a["move"] == b["move"] and a["stats"]["topMoves"] == b["stats"]["topMoves"]
        st_ = a["stats"]
        assert {"timeBudgetMs", "timeSpentMs", "nodesEvaluated", "maxDepthReached", "topMoves"} <= set(st_)
        assert st_["nodesEvaluated"] == 1000 and len(st_["topMoves"]) <= 10
        eng = DeviceUCT(125, n_trees=1, seed=7, max_depth=1, ucb_variant=1, reward_mode=1, leaf_batch=8)
        res = eng.search(dev.to_device(np.atleast_1d(st)), 125)
        assert int(res["visits"][0].sum()) == 1000 and int(res["counts"][0]) == len(legal)


# ---- transposition table across the moves of ONE agent (mcts_agent.py:337-339, 418-434) -----------------------------------
def test_agent_reuses_its_transposition_table_across_moves(golden_dir):
    """The second select_action of an agent -- on the position after its first move -- hits what the first search
    cached, reports the hits in stats["transposition_hits"], and equals the oracle run with ONE persistent table whose
    second search draws from the agent's next stream block."""
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    board = Board.from_packed(z["states"].view(STATE_DTYPE).reshape(-1)[74])      # 48 moves played, 16 legal moves: the
    gen = get_shared_generator()                                                  # second search re-visits the first one's nodes
    agent = MCTSAgent(iterations=160, seed=6, rollout_agent=RandomAgent(seed=6))
    ott = orc.OracleTT(18)
    zt = orc.zobrist_table(6)
    total_hits = 0
    for k in range(3):
        player = board.current_player
        legal = gen.get_legal_moves(board, player)
        r = agent.search_root(board, player)
        want = orc.mcts_search(board.to_packed(), 160, seed=6, tree_id=k, tt=ott, zobrist=zt, tt_namespace=0)
        n = want["n_root"]
        assert r["n_root"] == n == len(legal)
        assert np.array_equal(r["visits"], want["visits"][:n]), k
        assert np.array_equal(r["reward"], want["reward"][:n]), k
        assert agent.stats["transposition_hits"] == ott.hits(), k
        assert len(agent.transposition_table) == ott.size()
        if k:
            assert agent.stats["transposition_hits"] > total_hits, "a later search must find the earlier searches' entries"
        total_hits = agent.stats["transposition_hits"]
        best = int(np.argmax(r["visits"]))
        ids = sorted(m.to_id() for m in legal)
        m = Move.from_id(ids[best])
        assert board.place_piece(m.get_positions(gen.piece_orientations_cache[m.piece_id]), player, m.piece_id)
    st = agent.get_action_info()["transposition_stats"]
    assert st["hit_count"] == ott.hits() and st["access_count"] == ott.probes()
    ott.close()


def test_table_namespaces_isolate_concurrent_trees_and_persist(golden_dir):
    """Tree k of a launch uses namespace tt_tree_base + k whatever its tree_offset: concurrently searched trees do not
    see each other's entries, and a second launch with other stream ids finds the first launch's entries."""
    root_np = _roots(golden_dir)[3]
    root = dev.to_device(np.atleast_1d(root_np))
    zob = ZobristHash(seed=3)
    tt = TranspositionTable(max_size=1 << 16)
    otts = [orc.OracleTT(18) for _ in range(3)]             # the oracle runs the trees one by one: one table per namespace
    for launch, offset in enumerate((0, 40)):
        res = DeviceUCT(90, n_trees=3, seed=2, tree_offset=offset, tt=tt.device, zobrist=zob).search(root, 90)
        for k in range(3):
            want = orc.mcts_search(root_np, 90, seed=2, tree_id=offset + k, tt=otts[k], zobrist=zob.table, tt_namespace=k)
            n = want["n_root"]
            assert np.array_equal(res["visits"][k, :n].cpu().numpy(), want["visits"][:n]), (launch, k)
            assert np.array_equal(res["reward"][k, :n].cpu().numpy(), want["reward"][:n]), (launch, k)
    st = tt.device.stats()
    assert st["hits"] == sum(o.hits() for o in otts) > 0
    assert st["size"] == sum(o.size() for o in otts)
    for o in otts:
        o.close()


def test_transposition_table_evicts_oldest_tenth_when_full():
    """mcts/zobrist.py:196-200: a put into a full table first removes the max_size // 10 entries inserted first."""
    tt = TranspositionTable(max_size=50)
    ott = orc.OracleTT(8)
    keys = [int(k) for k in np.random.default_rng(0).integers(1, 2**63, size=120)]
    for i, k in enumerate(keys):
        if ott.size() >= 50:
            ott.evict_oldest(5)
        ott.put(k, float(i))
        tt.put(k, {"reward": float(i)})
        if i == 30:                                   # overwriting an entry does not renew it (dict order)
            ott.put(keys[2], -1.0)
            tt.put(keys[2], {"reward": -1.0})
    assert len(tt) == ott.size() <= 50
    for k in keys:
        got, want = tt.get(k), ott.get(k)
        assert (got is None) == (want is None)
        if got is not None:
            assert got["reward"] == want
    assert tt.get(keys[0]) is None and tt.get(keys[-1]) is not None
    ott.close()


# ---- the kernel against the REFERENCE's own MCTSAgent (tests/golden/mcts.npz) ----------------------------------------------
def _frontier_words(fr_bytes):
    bits = np.zeros((4, 448), dtype=np.uint8)
    bits[:, :400] = fr_bytes
    return torch.from_numpy(np.packbits(bits.reshape(4, 7, 64), axis=-1, bitorder="little").view("<i8").reshape(1, 4, 7).copy()).cuda()


def test_kernel_matches_reference_uct(golden_dir):
    """Every search the unmodified reference recorded (uniform / HeuristicAgent rollouts, transposition table and a second
    search of the same agent, dummy and fitted learned leaf evaluation, progressive bias, potential shaping, deep
    late-game trees): the kernel returns the same expanded moves, visits, reward sums, best move, hits and table sizes."""
    import uct_golden as U

    z, names, seed = U.load(golden_dir)
    zob = ZobristHash(seed=seed)
    model = torch.from_numpy(z["model"].copy())
    tables = {}
    for tag in names:
        sp = U.spec(tag)
        name = tag.split("/")[1]
        st, fr = U.root(z, tag)
        root = dev.to_device(np.atleast_1d(st))
        tt = None
        if sp.get("tt"):
            if sp.get("second"):
                tt = tables[name]                     # the table of the first search of this root (tt1/<name>)
            else:
                tt = tables[name] = TranspositionTable(max_size=1 << 15)
        kw = dict(n_trees=1, seed=seed, tree_offset=1 if sp.get("second") else 0, tt=tt.device if tt is not None else None,
                  zobrist=zob if tt is not None else None)              # (an empty table is falsy: it has __len__)
        before = tt.device.stats() if tt is not None else None
        if sp["family"] == "plain":
            eng = DeviceUCT(sp["iterations"], rollout_policy=sp.get("rollout_policy", 0), reward_mode=sp.get("reward_mode", 0), **kw)
            res = eng.search(root, sp["iterations"])
        elif sp["family"] == "learned":
            kind = sp.get("model_kind", 0)
            eng = DeviceUCT(sp["iterations"], reward_mode=3, leaf_model=torch.from_numpy(z["gbt_model"].copy()) if kind else model,
                            leaf_model_kind=kind, progressive_bias_weight=sp.get("bias_weight", 0.0), **kw)
            res = eng.search(root, sp["iterations"], root_frontiers=_frontier_words(fr))
        else:
            eng = DeviceUCT(sp["iterations"], reward_mode=4, leaf_model=model, max_rollout_moves=6, shaping_weight=2.0,
                            shaping_gamma=0.9, potential_mode="logit" if sp["potential_mode"] else "prob", **kw)
            res = eng.search(root, sp["iterations"], root_frontiers=_frontier_words(fr))
        assert int(res["status"][0]) == 0
        n = int(res["counts"][0])
        U.compare(z, tag, n, res["moves"][0].cpu().numpy(), res["visits"][0].cpu().numpy(), res["reward"][0].cpu().numpy(),
                  scale=sp.get("scale", 1.0))
        if tt is not None:
            after = tt.device.stats()
            want_hits, want_size = z[tag + "/tt"].tolist()
            first = 0 if not sp.get("second") else int(z["tt1/" + name + "/tt"][0])
            assert after["hits"] - before["hits"] == want_hits - first, tag
            assert after["size"] == want_size, tag


def test_device_table_eviction_matches_reference(golden_dir):
    """The same 163 puts of the reference's TranspositionTable(max_size=50) through the device table (mcts.TranspositionTable:
    evict_if_full + bk_tt_put): same keys and values left."""
    import json

    with open(os.path.join(golden_dir, "tt_eviction.json")) as f:
        fx = json.load(f)
    tt = TranspositionTable(max_size=fx["max_size"])
    for key, reward in fx["ops"]:
        tt.put(int(key), {"reward": reward})
    assert len(tt) == len(fx["left"])
    for key, reward in fx["left"].items():
        assert tt.get(int(key)) == {"reward": reward}
    gone = [k for k, _ in fx["ops"] if str(k) not in fx["left"]]
    assert gone and all(tt.get(int(k)) is None for k in gone[:40])

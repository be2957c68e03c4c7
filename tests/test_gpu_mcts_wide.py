"""K6: the round-synchronous (one tree over the whole GPU) search against its sequential restatement in the oracle."""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.mcts import DeviceUCT, WideUCT  # noqa: E402
from reinforcementlearning_blokus_b200.state import STATE_DTYPE, empty_states  # noqa: E402


def _roots(golden_dir):
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    states = z["states"].view(STATE_DTYPE).reshape(-1)
    return [empty_states(1)[0], states[3], states[54], states[74], states[42]]


def _same(res, want, k=0):
    n = want["n_root"]
    assert int(res["counts"][k]) == n
    assert int(res["status"][k]) == 0
    assert np.array_equal(res["visits"][k, :n].cpu().numpy(), want["visits"][:n])
    assert np.array_equal(res["reward"][k, :n].cpu().numpy(), want["reward"][:n])
    assert np.array_equal(res["moves"][k, :n].cpu().numpy(), want["moves"][:n])


@pytest.mark.parametrize("variant", ["flat", "uct", "const"])
def test_wide_search_matches_oracle(golden_dir, variant):
    kw = {"flat": dict(max_depth=1, ucb_variant=1, reward_mode=1), "uct": dict(), "const": dict(reward_mode=2)}[variant]
    for ri, root in enumerate(_roots(golden_dir)):
        for width, sims in ((48, 300), (256, 700)):
            want = orc.mcts_wide(root, [sims], width, seed=11 + ri, tree_id=5, **kw)
            eng = WideUCT(sims, width=width, seed=11 + ri, tree_offset=5, **kw)
            res = eng.search(dev.to_device(np.atleast_1d(root)), sims)
            _same(res, want)
            assert int(res["sims_done"][0]) == want["sims"] == sims
            assert int(res["visits"][0].sum()) == sims or want["n_root"] == 0


def test_wide_search_heuristic_rollouts_and_short_rollouts(golden_dir):
    for root in _roots(golden_dir)[1:4]:
        want = orc.mcts_wide(root, [200], 64, seed=3, rollout_policy=1)
        res = WideUCT(200, width=64, seed=3, rollout_policy=1).search(dev.to_device(np.atleast_1d(root)), 200)
        _same(res, want)
        want = orc.mcts_wide(root, [200], 64, seed=3, max_rollout_moves=5, exploration=0.7)
        res = WideUCT(200, width=64, seed=3, max_rollout_moves=5, exploration=0.7).search(dev.to_device(np.atleast_1d(root)), 200)
        _same(res, want)


def test_wide_full_tree_rounds_dealt_at_the_root(golden_dir):
    """Full trees with rounds of at least 128 walks: a fully expanded root deals the round's walks out to its children at once
    (threshold search, exploration scale frozen for the round) and every walk goes on below its child -- restated by the
    oracle as the greedy choice.  Wide rounds, resumed calls, uniform and HeuristicAgent rollouts, short rollouts."""
    for ri, root in enumerate(_roots(golden_dir)[:4]):
        for kw in (dict(max_rollout_moves=6), dict(max_rollout_moves=4, rollout_policy=1), dict(reward_mode=2)):
            calls = [700, 640] if ri % 2 else [1340]
            want = orc.mcts_wide(root, calls, 512, seed=21 + ri, tree_id=2, **kw)
            eng = WideUCT(sum(calls), width=512, seed=21 + ri, tree_offset=2, **kw)
            res = None
            for ci, c in enumerate(calls):
                res = eng.search(dev.to_device(np.atleast_1d(root)) if ci == 0 else None, c, resume=ci > 0)
            _same(res, want)
            assert int(res["sims_done"][0]) == want["sims"] == sum(calls)


def test_wide_search_resumes_and_runs_several_trees(golden_dir):
    roots_np = _roots(golden_dir)[:3]
    roots = dev.to_device(np.stack(roots_np))
    eng = WideUCT(600, width=128, n_trees=3, seed=9, tree_offset=100)
    eng.search(roots, 256)
    res = eng.search(None, 344, resume=True)
    for k in range(3):
        want = orc.mcts_wide(roots_np[k], [256, 344], 128, seed=9, tree_id=100 + k)
        _same(res, want, k)
    # a second engine with the same arguments gives the same answer (no dependence on scheduling)
    eng2 = WideUCT(600, width=128, n_trees=3, seed=9, tree_offset=100)
    eng2.search(roots, 256)
    res2 = eng2.search(None, 344, resume=True)
    assert torch.equal(res["visits"], res2["visits"]) and torch.equal(res["reward"], res2["reward"])
    with pytest.raises(ValueError):
        eng.search(None, 1, resume=True)


def test_width_one_is_the_sequential_search(golden_dir):
    """With one simulation per round there is nothing pending when a walk starts: the schedule degenerates to the
    reference's iteration -- same tree as k_mcts up to the rollout streams, so compare the structure with the oracle's
    sequential UCT in the constant-leaf mode (no rollouts: rewards are fixed, only selection/expansion decide)."""
    root = _roots(golden_dir)[3]
    want = orc.mcts_search(root, 150, seed=1, reward_mode=2)
    res = WideUCT(150, width=1, seed=1, reward_mode=2).search(dev.to_device(np.atleast_1d(root)), 150)
    _same(res, want)
    seq = DeviceUCT(150, n_trees=1, seed=1, reward_mode=2).search(dev.to_device(np.atleast_1d(root)), 150)
    assert torch.equal(seq["visits"][0], res["visits"][0])


def test_agents_use_the_wide_search(golden_dir):
    """FastMCTSAgent.think runs the flat wide search by default; MCTSAgent(search_width=...) the full-tree one."""
    from reinforcementlearning_blokus_b200 import Board, FastMCTSAgent, MCTSAgent, RandomAgent, get_shared_generator

    board = Board.from_packed(_roots(golden_dir)[2])
    player = board.current_player
    legal = get_shared_generator().get_legal_moves(board, player)
    ids = sorted(m.to_id() for m in legal)
    st = board.to_packed()
    # think(): the budget is split over root-parallel trees (one planner each), width = a quarter of a tree's share;
    # the root statistics of the trees are summed
    from reinforcementlearning_blokus_b200 import Move
    n = len(ids)
    for budget in (400, 1000, 5000):
        agent = FastMCTSAgent(iterations=budget, time_limit=30.0, seed=5)
        out = agent.think(board, player, legal, 30000)
        per_tree, width, trees = agent._wide_cache[0][:3]
        assert trees == max(1, min(8, (budget + 1250) // 2500)) and per_tree * trees >= budget
        tot = np.zeros(n, dtype=np.int64)
        chunk = max(FastMCTSAgent.CHUNK, 4 * width)             # think() checks the clock between chunks of four rounds
        calls = [min(chunk, per_tree - lo) for lo in range(0, per_tree, chunk)]
        for k in range(trees):
            w = orc.mcts_wide(st, calls, width, seed=5, tree_id=k, max_depth=1, ucb_variant=1, reward_mode=1)
            assert w["n_root"] == n
            tot += w["visits"][:n]
        assert out["stats"]["nodesEvaluated"] == per_tree * trees == int(tot.sum())
        assert out["move"].to_id() == ids[max(range(n - 1, -1, -1), key=lambda r: tot[r])]
        for t in out["stats"]["topMoves"]:
            assert t["visits"] == tot[ids.index(Move(t["piece_id"], t["orientation"], t["anchor_row"], t["anchor_col"]).to_id())]
    # the sequential search stays available
    seq = FastMCTSAgent(iterations=200, time_limit=30.0, seed=5, search_width=0).think(board, player, legal, 30000)
    w2 = orc.mcts_search(st, 200, seed=5, tree_id=0, max_depth=1, ucb_variant=1, reward_mode=1)
    assert seq["move"].to_id() == ids[max(range(n - 1, -1, -1), key=lambda r: w2["visits"][r])]
    # MCTSAgent with the wide full-tree search, heuristic rollouts (its default policy) and uniform ones
    for rollout, pol in ((None, 1), (RandomAgent(seed=1), 0)):
        agent = MCTSAgent(iterations=400, seed=8, search_width=128, rollout_agent=rollout, use_transposition_table=False)
        r = agent.search_root(board, player)
        w3 = orc.mcts_wide(st, [400], 128, seed=8, tree_id=0, rollout_policy=pol)
        assert np.array_equal(r["visits"], w3["visits"][:n]) and np.array_equal(r["reward"], w3["reward"][:n])
        assert agent.select_action(board, player, legal) in legal


def test_wide_search_degenerate_roots(golden_dir):
    """A root whose player cannot move (every simulation re-simulates the root) and roots with very few moves (the
    tree is exhausted early and later walks keep landing on terminal nodes)."""
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    states = z["states"].view(STATE_DTYPE).reshape(-1)
    picked = {}
    for s in states:
        n = len(orc.OracleBoard.from_packed(s).legal_moves(int(s["to_move"])))
        if n <= 3 and n not in picked:
            picked[n] = s
    assert 0 in picked and len(picked) >= 2
    for n, root in sorted(picked.items()):
        for kw in (dict(), dict(max_depth=1, ucb_variant=1, reward_mode=1)):
            want = orc.mcts_wide(root, [150], 64, seed=4, **kw)
            res = WideUCT(150, width=64, seed=4, **kw).search(dev.to_device(np.atleast_1d(root)), 150)
            assert int(res["counts"][0]) == want["n_root"] == n
            assert int(res["sims_done"][0]) == 150 and int(res["status"][0]) == 0
            _same(res, want)

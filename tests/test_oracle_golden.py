"""Pins the CPU oracle (oracle/blokus_oracle.c) against outputs of the Python reference.

The fixtures under tests/golden/ were produced by tests/golden/generate_golden.py, which imports
the unmodified reference.  These tests run without a GPU.
"""

import hashlib
import json
import os

import numpy as np
import pytest

from oracle import orc


def _load_json(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def _state_from_hex(h):
    return np.frombuffer(bytes.fromhex(h), dtype=orc.STATE_DTYPE)[0]


def test_tables_match_reference(golden_dir):
    ref = _load_json(golden_dir, "tables.json")["orientations"]
    mine = orc.orientation_table()
    assert len(mine) == 91
    for a, b in zip(ref, mine):
        assert a["piece_id"] == b["piece_id"] and a["index"] == b["index"]
        assert [tuple(c) for c in a["cells"]] == b["cells"]
    # 88 squares, S == Z
    sizes = {}
    for o in mine:
        sizes[o["piece_id"]] = len(o["cells"])
    assert sum(sizes.values()) == 88
    assert sum(len(o["cells"]) for o in mine) == 410
    s = sorted(tuple(o["cells"]) for o in mine if o["piece_id"] == 9)
    z = sorted(tuple(o["cells"]) for o in mine if o["piece_id"] == 10)
    assert s == z


def test_python_tables_match_reference(golden_dir):
    from reinforcementlearning_blokus_b200 import pieces

    ref = _load_json(golden_dir, "tables.json")["orientations"]
    for a, o in zip(ref, pieces.ORIENTATIONS):
        assert (a["piece_id"], a["index"]) == (o.piece_id, o.index)
        assert [tuple(c) for c in a["cells"]] == list(o.cells)
    counts = [len(pieces.PIECE_ORIENTATIONS[p]) for p in range(1, 22)]
    assert counts == [1, 2, 2, 4, 2, 1, 4, 8, 4, 4, 8, 2, 8, 8, 8, 4, 4, 4, 4, 1, 8]


def test_first_move_counts():
    b = orc.OracleBoard()
    for p in range(4):
        assert len(b.legal_moves(p)) == 58
        assert len(b.legal_moves(p, naive=True)) == 58


@pytest.mark.parametrize("name", ["kat1_states.json", "kat2_game.json"])
def test_kat_states(golden_dir, name):
    data = _load_json(golden_dir, name)
    if isinstance(data, dict):
        data = [data]
    for d in data:
        b = orc.OracleBoard.from_packed(_state_from_hex(d["packed_hex"]))
        assert b.grid.astype(int).tolist() == d["grid"]
        assert b.current_player == d["to_move"]
        assert b.move_count == d["move_count"]
        for p in range(4):
            want = np.array(d["legal"][p], dtype=np.uint16)
            assert np.array_equal(b.legal_moves(p), want)
            assert np.array_equal(b.legal_moves(p, naive=True), want)
            assert b.has_legal_moves(p) == d["has_legal_moves"][p]
            assert b.board_score(p) == d["board_score"][p]
            assert b.game_score(p) == d["game_score"][p]
            # recomputed frontier is the true frontier: a subset of the reference's incremental set (Q8)
            assert set(b.frontier(p)) <= {tuple(x) for x in d["frontier"][p]}


def test_kat2_replay_incremental(golden_dir):
    """Replays the reference's recorded 20-move game move by move (tests/fixtures/advanced_metrics_snapshot.json)."""
    d = _load_json(golden_dir, "kat2_game.json")
    b = orc.OracleBoard()
    for mv in d["moves"]:
        assert b.current_player == mv["player"]
        assert b.is_move_legal(mv["player"], mv["move_id"])
        assert mv["move_id"] in b.legal_moves(mv["player"])
        assert b.make_move(mv["player"], mv["move_id"])
        assert [b.game_score(p) for p in range(4)] == mv["game_score"]
    assert [b.game_score(p) for p in range(4)] == [27, 27, 23, 26]
    # incremental frontier == the reference's incremental frontier, exactly
    for p in range(4):
        assert b.frontier(p) == [tuple(x) for x in d["frontier"][p]]
    assert b.pack().tobytes().hex() == d["packed_hex"]


def test_random_states(golden_dir):
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    states = z["states"].view(orc.STATE_DTYPE).reshape(-1)
    offs, ids = z["offsets"], z["move_ids"]
    for i, s in enumerate(states):
        b = orc.OracleBoard.from_packed(s)
        assert np.array_equal(b.grid, z["grids"][i])
        for p in range(4):
            want = ids[offs[4 * i + p]:offs[4 * i + p + 1]]
            got = b.legal_moves(p)
            assert np.array_equal(got, want), (i, p)
            if i % 8 == 0:
                assert np.array_equal(b.legal_moves(p, naive=True), want)
            assert b.has_legal_moves(p) == bool(z["has_moves"][i, p])
            assert b.board_score(p) == z["board_scores"][i, p]
            assert b.game_score(p) == z["game_scores"][i, p]
            for mid in want[:: max(1, len(want) // 5)]:
                assert b.is_move_legal(p, mid)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert orc.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert orc.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_full_games_match_reference(golden_dir):
    """Philox-indexed games replayed through the reference's BlokusGame give the same traces and scores."""
    z = np.load(os.path.join(golden_dir, "games.npz"))
    seed = int(z["seed"][0])
    for gi in range(z["traces"].shape[0]):
        for fast in (True, False) if gi < 3 else (True,):
            r = orc.playout(seed, gi, fast=fast, want_final=True)
            assert np.array_equal(r["trace"], z["traces"][gi]), gi
            assert np.array_equal(r["choice"], z["choices"][gi])
            assert np.array_equal(r["game_scores"], z["game_scores"][gi])
            assert np.array_equal(r["board_scores"], z["board_scores"][gi])
            assert r["turns"] == z["turns"][gi]
            assert r["final"].tobytes() == z["finals"][gi].tobytes()


def test_zobrist_matches_reference(golden_dir):
    d = _load_json(golden_dir, "zobrist.json")
    boards = [orc.OracleBoard.from_packed(_state_from_hex(h)) for h in d["states_hex"]]
    for seed in (0, 42, 12345):
        t = orc.zobrist_table(seed)
        g = d[str(seed)]
        assert hashlib.sha256(t.astype("<u8").tobytes()).hexdigest() == g["table_sha256"]
        assert [int(x) for x in t[:8]] == g["first"]
        assert [int(x) for x in t[2000:2004]] == g["turn"]
        assert int(t[2004]) == g["piece00"]
        assert [b.zobrist_hash(t) for b in boards] == g["hashes"]
        # numpy's MT19937 stream is what the reference draws from (mcts/zobrist.py:33)
        rs = np.random.RandomState(seed)
        assert int(rs.randint(0, 2**64, dtype=np.uint64)) == int(t[0])
    assert d["0"]["hashes"][0] == 0x72F88A837C756705


def test_heuristic_scores_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "heuristic.npz"))
    states = z["states"].view(orc.STATE_DTYPE).reshape(-1)
    offs = z["offsets"]
    for i, s in enumerate(states):
        b = orc.OracleBoard.from_packed(s)
        p = int(s["to_move"])
        ids = z["move_ids"][offs[i]:offs[i + 1]]
        assert np.array_equal(b.legal_moves(p), ids)
        got = np.array([b.heuristic_score(p, m) for m in ids])
        # same float64 operations in the same order -> bit-identical
        assert np.array_equal(got, z["scores"][offs[i]:offs[i + 1]])


# ---- snapshot feature rows -------------------------------------------------------------
def test_snapshot_features_of_replayed_games_match_reference(golden_dir):
    """features.py rows captured by the reference at the arena checkpoints of the 16 golden games: every float equal."""
    z = np.load(os.path.join(golden_dir, "snapshots.npz"))
    g = np.load(os.path.join(golden_dir, "games.npz"))
    ck, mt = z["checkpoints"], int(z["max_turns"][0])
    fronts = np.unpackbits(z["frontiers"], axis=-1, bitorder="little")[..., :400]
    for gi in range(g["traces"].shape[0]):
        r = orc.replay_snapshots(g["traces"][gi], ck, mt)
        assert (r["turn_index"] == z["turn_index"][gi]).all()
        hit = r["turn_index"] >= 0
        assert hit.sum() >= 8
        assert (r["frontiers"][hit] == fronts[gi][hit]).all()            # incremental (stale cells included) sets
        assert (r["states"].view(np.uint8).reshape(len(ck), 256)[hit] == z["snap_states"][gi][hit]).all()
        assert (r["features"][hit] == z["features"][gi][hit]).all()


def test_snapshot_features_of_state_corpus_match_reference(golden_dir):
    """The learned-evaluator call (turn_index = move_count, rebuilt frontier sets) on the 240-state corpus."""
    z = np.load(os.path.join(golden_dir, "snapshots.npz"))
    states = np.load(os.path.join(golden_dir, "random_states.npz"))["states"].view(orc.STATE_DTYPE).reshape(-1)
    for i in range(len(states)):
        assert (orc.snapshot_features(states[i]) == z["state_features"][i]).all(), i


# ---- UCT: the oracle's restatement against the reference's own MCTSAgent -------------------------------------------------
def test_uct_matches_reference(golden_dir):
    """mcts/mcts_agent.py:19-191, 304-582 executed by the reference (tests/golden/mcts.npz: uniform and HeuristicAgent
    rollouts, transposition table incl. a second search of the same agent, dummy / fitted learned leaf evaluation,
    progressive bias, potential shaping, deep late-game trees): identical expanded moves, visits, reward sums, best
    move, transposition hits and table sizes."""
    import uct_golden as U

    z, names, seed = U.load(golden_dir)
    zt = orc.zobrist_table(seed)
    model = z["model"]
    tables = {}
    assert len(names) >= 48
    for tag in names:
        sp = U.spec(tag)
        name = tag.split("/")[1]
        st, fr = U.root(z, tag)
        tt = None
        if sp.get("tt"):
            if sp.get("second"):
                tt = tables[name]                                   # the first search's table (tt1/<name> ran before)
            else:
                tt = tables[name] = orc.OracleTT(16)
        kw = dict(tree_id=1 if sp.get("second") else 0, tt=tt, zobrist=zt if tt else None)
        if sp["family"] == "plain":
            w = orc.mcts_search(st, sp["iterations"], seed=seed, rollout_policy=sp.get("rollout_policy", 0),
                                reward_mode=sp.get("reward_mode", 0), **kw)
        elif sp["family"] == "learned":
            kind = sp.get("model_kind", 0)
            w = orc.mcts_search_learned(st, sp["iterations"], seed=seed, model=z["gbt_model"] if kind else model, root_frontier=fr,
                                        bias_weight=sp.get("bias_weight", 0.0), model_kind=kind, **kw)
        else:
            w = orc.mcts_search_shaped(st, sp["iterations"], seed=seed, model=model, root_frontier=fr, max_rollout_moves=6,
                                       shaping_weight=2.0, shaping_gamma=0.9, potential_mode=sp["potential_mode"], **kw)
        U.compare(z, tag, w["n_root"], w["moves"], w["visits"], w["reward"], scale=sp.get("scale", 1.0))
        if tt is not None:
            assert [tt.hits(), tt.size()] == z[tag + "/tt"].tolist(), tag
    for t in tables.values():
        t.close()


def test_heuristic_policy_games_match_reference(golden_dir):
    """Whole games played by the reference's HeuristicAgent.select_action (its scores, its softmax; only the categorical
    draw is Philox, tests/golden/generate_golden.py gen_heuristic_games): the oracle's heuristic-policy playout picks the
    same ~430 moves and ends on the same scores."""
    z = np.load(os.path.join(golden_dir, "heuristic_games.npz"))
    seed = int(z["seed"][0])
    for g in range(len(z["traces"])):
        o = orc.playout(seed, g, policy=1)
        assert np.array_equal(o["trace"], z["traces"][g]), g
        assert np.array_equal(o["game_scores"], z["game_scores"][g]) and o["turns"] == int(z["turns"][g])


def test_transposition_table_eviction_matches_reference(golden_dir):
    """mcts/zobrist.py:188-202 run by the reference on 163 puts into a 50-entry table (tests/golden/tt_eviction.json): the
    oracle's table -- evict the max_size // 10 entries inserted first when full, overwriting keeps an entry's place --
    ends with the same keys and values."""
    with open(os.path.join(golden_dir, "tt_eviction.json")) as f:
        fx = json.load(f)
    t = orc.OracleTT(8)
    for key, reward in fx["ops"]:
        if t.size() >= fx["max_size"]:
            t.evict_oldest(fx["max_size"] // 10)
        t.put(int(key), reward)
    assert t.size() == len(fx["left"])
    for key, reward in fx["left"].items():
        assert t.get(int(key)) == reward
    for key, _ in fx["ops"]:
        if str(key) not in fx["left"]:
            assert t.get(int(key)) is None
    t.close()

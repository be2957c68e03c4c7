"""GPU parity tests: CUDA kernels (through the C ABI) vs the pinned oracle and the reference goldens.

Bit-exact everywhere: move ids, counts, traces, scores and packed states are integers."""

import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.state import STATE_DTYPE, empty_states  # noqa: E402


def _np(t):
    return t.cpu().numpy()


def _states_to_dev(states_np):
    return dev.to_device(states_np)


@pytest.fixture(scope="module")
def golden_states(golden_dir):
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def oracle_corpus():
    """~1500 states sampled along oracle-played random games (all stages incl. terminal)."""
    rng = np.random.RandomState(7)
    states = []
    for g in range(60):
        r = orc.playout(1234, g)
        b = orc.OracleBoard()
        cut = set(rng.choice(r["turns"] + 1, size=min(25, r["turns"] + 1), replace=False).tolist())
        for t in range(r["turns"] + 1):
            if t in cut:
                states.append(b.pack())
            if t == r["turns"]:
                break
            mid = int(r["trace"][t])
            if mid == orc.PASS_ID:
                b.pass_turn()
            else:
                assert b.make_move(b.current_player, mid)
    return np.array(states, dtype=STATE_DTYPE)


def _csr_lists(counts, offsets, ids):
    counts, offsets, ids = _np(counts), _np(offsets), _np(ids)
    assert offsets[0] == 0 and np.array_equal(np.diff(offsets.astype(np.int64)), counts)
    return [ids[offsets[i]:offsets[i + 1]] for i in range(len(counts))]


def test_movegen_matches_reference_goldens(golden_states):
    states = golden_states["states"]
    n = states.shape[0]
    d = _states_to_dev(states)
    offs, ids = golden_states["offsets"], golden_states["move_ids"]
    for p in range(4):
        players = torch.full((n,), p, dtype=torch.uint8, device="cuda")
        lists = _csr_lists(*dev.movegen(d, players))
        bitmap = _np(dev.movegen_bitmap(d, players))
        for i in range(n):
            want = ids[offs[4 * i + p]:offs[4 * i + p + 1]]
            assert np.array_equal(lists[i], want), (i, p)
            # bitmap <-> id list
            got = []
            for w in np.nonzero(bitmap[i])[0]:
                v = int(bitmap[i][w]) & 0xFFFFF
                got += [int(w) * 20 + c for c in range(20) if v >> c & 1]
            assert got == want.tolist()
    hm = _np(dev.has_moves(d))
    want_mask = (golden_states["has_moves"] * np.array([1, 2, 4, 8], dtype=np.uint8)).sum(axis=1)
    assert np.array_equal(hm, want_mask.astype(np.uint8))
    gs, bs = dev.scores(d)
    assert np.array_equal(_np(gs), golden_states["game_scores"])
    assert np.array_equal(_np(bs), golden_states["board_scores"])


def test_movegen_side_to_move_matches_oracle(oracle_corpus):
    d = _states_to_dev(oracle_corpus)
    lists = _csr_lists(*dev.movegen(d))
    counts = _np(dev.movegen_count(d))
    for i, s in enumerate(oracle_corpus):
        b = orc.OracleBoard.from_packed(s)
        want = b.legal_moves(int(s["to_move"]))
        assert np.array_equal(lists[i], want), i
        assert counts[i] == len(want)


def test_kat1_benchmark_states(golden_dir):
    with open(os.path.join(golden_dir, "kat1_states.json")) as f:
        kat = json.load(f)
    states = np.array([np.frombuffer(bytes.fromhex(k["packed_hex"]), dtype=STATE_DTYPE)[0] for k in kat])
    d = _states_to_dev(states)
    for p in range(4):
        players = torch.full((len(kat),), p, dtype=torch.uint8, device="cuda")
        lists = _csr_lists(*dev.movegen(d, players))
        for i, k in enumerate(kat):
            assert lists[i].tolist() == k["legal"][p]
    # SURVEY 8c KAT-1 counts
    assert [len(k["legal"][0]) for k in kat] == [279, 476, 288, 407, 65]


def test_first_move_is_58_for_everyone():
    s = empty_states(4)
    d = _states_to_dev(s)
    players = torch.arange(4, dtype=torch.uint8, device="cuda")
    assert _np(dev.movegen_count(d, players)).tolist() == [58, 58, 58, 58]


def test_is_move_legal_and_step(golden_dir, oracle_corpus):
    rng = np.random.RandomState(3)
    sub = oracle_corpus[::7]
    n = len(sub)
    d = _states_to_dev(sub)
    # random ids: legality must agree with the oracle's is_move_legal (engine/move_generator.py:912)
    for rep in range(3):
        ids = rng.randint(0, 36400, size=n).astype(np.uint16)
        if rep == 2:
            ids[:] = 40000  # out of range -> illegal, never a fault
        got = _np(dev.is_move_legal(d, torch.from_numpy(ids).cuda()))
        for i, s in enumerate(sub):
            b = orc.OracleBoard.from_packed(s)
            assert bool(got[i]) == b.is_move_legal(int(s["to_move"]), int(ids[i])), (rep, i)
    # one legal move each: step must produce the oracle's next state byte for byte
    ids = np.full(n, orc.PASS_ID, dtype=np.uint16)
    want_states = []
    want_alive = []
    for i, s in enumerate(sub):
        b = orc.OracleBoard.from_packed(s)
        legal = b.legal_moves(int(s["to_move"]))
        if len(legal):
            ids[i] = legal[rng.randint(len(legal))]
            assert b.make_move(int(s["to_move"]), int(ids[i]))
        else:
            b.pass_turn()
        want_states.append(b.pack())
        want_alive.append(sum((1 << p) for p in range(4) if b.has_legal_moves(p)))
    work = d.clone()
    ok, alive = dev.step(work, torch.from_numpy(ids).cuda(), want_alive=True)
    assert _np(ok).all()
    assert np.array_equal(_np(work), np.array(want_states, dtype=STATE_DTYPE).view(np.uint8).reshape(n, 256))
    assert _np(alive).tolist() == want_alive
    # an illegal move is rejected and leaves the state untouched (make_move returns False)
    bad = np.full(n, 91 * 400 - 1, dtype=np.uint16)
    work2 = d.clone()
    ok2, _ = dev.step(work2, torch.from_numpy(bad).cuda())
    assert not _np(ok2).any() and torch.equal(work2, d)
    # four passes bring every state back byte for byte (load/store round trip)
    work3 = d.clone()
    passes = torch.full((n,), orc.PASS_ID, dtype=torch.uint16, device="cuda")
    for _ in range(4):
        dev.step(work3, passes)
    assert torch.equal(work3, d)


def test_kat2_recorded_game_replays(golden_dir):
    with open(os.path.join(golden_dir, "kat2_game.json")) as f:
        kat = json.load(f)
    s = _states_to_dev(empty_states(1))
    for mv in kat["moves"]:
        ok, _ = dev.step(s, torch.tensor([mv["move_id"]], dtype=torch.uint16, device="cuda"))
        assert _np(ok)[0] == 1
        gs, _ = dev.scores(s)
        assert _np(gs)[0].tolist() == mv["game_score"]
    assert _np(s)[0].tobytes().hex() == kat["packed_hex"]


def test_playout_matches_reference_games(golden_dir):
    """Games generated by the Python reference + Philox indices (tests/golden/games.npz)."""
    z = np.load(os.path.join(golden_dir, "games.npz"))
    n = z["traces"].shape[0]
    r = dev.playout(n, seed=int(z["seed"][0]), want_trace=True, want_choice=True, want_final=True)
    assert np.array_equal(_np(r["trace"]), z["traces"])
    assert np.array_equal(_np(r["choice"]), z["choices"])
    assert np.array_equal(_np(r["game_scores"]), z["game_scores"])
    assert np.array_equal(_np(r["board_scores"]), z["board_scores"])
    assert np.array_equal(_np(r["final_states"]), z["finals"])


def test_heuristic_playouts_match_reference_games(golden_dir):
    """bk_playout(policy = HeuristicAgent) against whole games played by the reference's own HeuristicAgent
    (tests/golden/heuristic_games.npz): same moves, same final scores."""
    z = np.load(os.path.join(golden_dir, "heuristic_games.npz"))
    n = len(z["traces"])
    r = dev.playout(n, seed=int(z["seed"][0]), policy=dev.POLICY_HEURISTIC, want_trace=True)
    assert np.array_equal(_np(r["trace"]), z["traces"])
    assert np.array_equal(_np(r["game_scores"]), z["game_scores"])


def test_playout_matches_oracle_and_is_batch_independent():
    seed, n, off = 20261018, 256, 1000
    r = dev.playout(n, seed=seed, game_offset=off, want_trace=True, want_choice=True, want_final=True)
    tr, gs, bs, fin = _np(r["trace"]), _np(r["game_scores"]), _np(r["board_scores"]), _np(r["final_states"])
    for i in range(0, n, 4):
        o = orc.playout(seed, off + i, want_final=True)
        assert np.array_equal(tr[i], o["trace"]), i
        assert np.array_equal(gs[i], o["game_scores"]) and np.array_equal(bs[i], o["board_scores"])
        assert fin[i].tobytes() == o["final"].tobytes()
    # the same global game index gives the same game in any batch
    r2 = dev.playout(16, seed=seed, game_offset=off + 100, want_trace=True)
    assert np.array_equal(_np(r2["trace"]), tr[100:116])


def test_playout_from_start_states_and_max_turns(oracle_corpus):
    sub = oracle_corpus[::11]
    n = len(sub)
    d = _states_to_dev(sub)
    r = dev.playout(n, seed=5, game_offset=77, start_states=d, want_trace=True, want_final=True)
    rt = dev.playout(n, seed=5, game_offset=77, start_states=d, max_turns=6, want_trace=True, want_final=True)
    for i in range(0, n, 3):
        o = orc.playout(5, 77 + i, start=sub[i], want_final=True)
        assert np.array_equal(_np(r["trace"])[i], o["trace"])
        assert np.array_equal(_np(r["game_scores"])[i], o["game_scores"])
        assert _np(r["final_states"])[i].tobytes() == o["final"].tobytes()
        o6 = orc.playout(5, 77 + i, start=sub[i], max_turns=6, want_final=True)
        assert np.array_equal(_np(rt["trace"])[i], o6["trace"])
        assert _np(rt["final_states"])[i].tobytes() == o6["final"].tobytes()


def test_playout_full_batch_invariants():
    """BASELINE config 2 size: 65,536 concurrent games; size-independent properties on every game."""
    from reinforcementlearning_blokus_b200 import pieces

    n = 65536
    r = dev.playout(n, seed=42, want_trace=True, want_final=True)
    fin = _np(r["final_states"]).view(STATE_DTYPE).reshape(n)
    gs, bs, tr = _np(r["game_scores"]), _np(r["board_scores"]), _np(r["trace"])
    bits = fin["bits"]                                    # [n,4,7]
    pop = np.zeros((n, 4), dtype=np.int64)
    for w in range(7):
        pop += np.array([bin(int(x)).count("1") for x in bits[:, :, w].reshape(-1)]).reshape(n, 4)
    sizes = np.array([pieces.PIECE_SIZES[p] for p in range(1, 22)])
    used_cells = ((fin["used"][:, :, None] >> np.arange(21)) & 1) @ sizes
    assert np.array_equal(pop, used_cells)                # cells on board == sizes of spent pieces
    allp = (fin["used"] & 0x1FFFFF) == 0x1FFFFF
    assert np.array_equal(bs, pop + 15 * allp)            # Board.get_score identity
    assert (gs >= bs).all()
    # no overlap between players
    acc = np.zeros((n, 7), dtype=np.uint64)
    for p in range(4):
        assert not (acc & bits[:, p]).any()
        acc |= bits[:, p]
    # every game is over: nobody can move in the final state
    assert not _np(dev.has_moves(r["final_states"])).any()
    # placements in the trace == move_count; trace ids decode to spent pieces
    placed = (tr != 0xFFFF).sum(axis=1)
    assert np.array_equal(placed, fin["move_count"])
    assert 40 <= placed.mean() <= 80
    # a sample agrees with the oracle exactly
    for i in (0, 1, 777, 40000, 65535):
        o = orc.playout(42, i)
        assert np.array_equal(tr[i], o["trace"]) and np.array_equal(gs[i], o["game_scores"])


def test_zobrist_and_tt(golden_dir, oracle_corpus):
    with open(os.path.join(golden_dir, "zobrist.json")) as f:
        z = json.load(f)
    states = np.array([np.frombuffer(bytes.fromhex(h), dtype=STATE_DTYPE)[0] for h in z["states_hex"]])
    d = _states_to_dev(states)
    for seed in (0, 42, 12345):
        table = orc.zobrist_table(seed)
        rs = np.random.RandomState(seed)   # what the reference draws from (mcts/zobrist.py:33)
        assert int(rs.randint(0, 2**64, dtype=np.uint64)) == int(table[0])
        t = torch.from_numpy(table.view(np.int64)).cuda()
        keys = _np(dev.zobrist_hash(d, t)).view(np.uint64)
        assert [int(k) for k in keys] == z[str(seed)]["hashes"]
    # larger sample against the oracle's hash_board
    table = orc.zobrist_table(7)
    t = torch.from_numpy(table.view(np.int64)).cuda()
    dc = _states_to_dev(oracle_corpus)
    keys_t = dev.zobrist_hash(dc, t)
    keys = _np(keys_t).view(np.uint64)
    for i in range(0, len(oracle_corpus), 9):
        assert int(keys[i]) == orc.OracleBoard.from_packed(oracle_corpus[i]).zobrist_hash(table)
    # transposition table: put / get / miss / overwrite / clear
    tt = dev.DeviceTranspositionTable(log2_slots=12)
    uniq, first = np.unique(keys, return_index=True)
    k = torch.from_numpy(uniq.view(np.int64)).cuda()
    rewards = torch.arange(len(uniq), dtype=torch.float64, device="cuda") - 7.0
    got, found = tt.get(k)
    assert not _np(found).any()
    tt.put(k, rewards)
    got, found = tt.get(k)
    assert _np(found).all() and torch.equal(got, rewards)
    assert tt.stats()["size"] == len(uniq)
    tt.put(k[:5], rewards[:5] + 100.0)             # dict semantics: last put wins, size unchanged
    got, found = tt.get(k[:5])
    assert torch.equal(got, rewards[:5] + 100.0) and tt.stats()["size"] == len(uniq)
    miss = torch.tensor([1, 2, 3], dtype=torch.int64, device="cuda")
    _, f2 = tt.get(miss)
    assert not _np(f2).any()
    tt.clear()
    _, f3 = tt.get(k)
    assert not _np(f3).any() and tt.stats()["size"] == 0
    tt.close()


def test_exclusive_scan_sizes():
    g = torch.Generator(device="cpu").manual_seed(1)
    for n in (1, 2, 1023, 1024, 1025, 5000, (1 << 20) + 77, 3 * (1 << 20) + 5):
        c = torch.randint(0, 900, (n,), generator=g, dtype=torch.int32)
        got = dev.exclusive_scan(c.cuda()).cpu().numpy().astype(np.int64)
        want = np.concatenate([[0], np.cumsum(c.numpy().astype(np.int64))])
        assert np.array_equal(got, want), n


def test_legality_sweep_million():
    """BASELINE.json configs[2] at its stated size: 10^6 synthetic states (every game stage, terminal states included),
    the device's ascending move list of the side to move against the oracle's frontier+bitboard generator byte for
    byte, every 1000th state also against the naive generator (0.1 s of GPU time, ~10-20 s of oracle time on the host
    cores; tools/legality_sweep.py runs 10^7)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "legality_sweep", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "legality_sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = mod.sweep(1_000_000, seed=77)
    assert res["states"] == 1_000_000 and res["mismatches"] == 0 and res["naive_mismatches"] == 0
    assert res["terminal_states"] > 0 and res["max_moves"] > 800 and res["total_moves"] > 150_000_000
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        import json

        with open(os.path.join(out, "legality_sweep_1M.json"), "w") as f:
            json.dump(res, f)


def test_abi_error_codes_and_empty_batches():
    """Never throws, never faults: bad arguments come back as negative codes with a message; n = 0 is a no-op."""
    import ctypes as C

    from reinforcementlearning_blokus_b200 import _lib

    lib = _lib.init(torch.cuda.current_device())
    st = _states_to_dev(empty_states(2))
    counts = torch.zeros(2, dtype=torch.int32, device="cuda")
    # n = 0 everywhere
    assert lib.bk_movegen_count(None, None, 0, None, None) == 0
    assert lib.bk_playout(None, 0, 1, 0, 0, 0, 0, None, None, None, None, None, None) == 0
    assert lib.bk_scores(None, 0, None, None, None) == 0
    e = dev.movegen(st[:0].contiguous())
    assert e[0].numel() == 0 and e[1].tolist() == [0] and e[2].numel() == 0
    # bad arguments
    assert lib.bk_movegen_count(None, None, 2, counts.data_ptr(), None) == -2
    assert b"null" in lib.bk_last_error()
    assert lib.bk_playout(None, 4, 1, 0, 7, 0, 0, None, None, None, None, None, None) == -2          # unknown policy
    assert lib.bk_playout(None, 4, 1, 0, 0, 0, 500, None, None, None, None, None, None) == -2        # max_turns > 96
    assert lib.bk_playout(None, 4, 1, 0, 0, 50, 10, None, None, None, None, None, None) == -2        # first_turn > max_turns
    h = C.c_void_p()
    assert lib.bk_tt_create(C.byref(h), 2) == -2 and lib.bk_tt_create(C.byref(h), 40) == -2
    p = _lib.MctsParams(iterations=1, iterations_total=8, max_rollout_moves=5, rollout_policy=0, max_depth=0, ucb_variant=0,
                        resume=0, reward_mode=0, exploration=1.0, seed=1, tree_offset=0)
    ws = torch.zeros(16, dtype=torch.uint8, device="cuda")
    assert lib.bk_mcts_search(st.data_ptr(), 1, 1, C.byref(p), None, None, ws.data_ptr(), 16, 8, None, None, None, None,
                              None, None) == -4                                                       # workspace too small
    assert lib.bk_init(99) == -2
    # outputs a caller does not want may be NULL
    assert lib.bk_playout(None, 4, 1, 0, 0, 0, 0, None, None, None, None, None, None) == 0
    torch.cuda.synchronize()
    # a game that is already over stays where it is
    fin = dev.playout(8, seed=3, want_final=True)["final_states"]
    again = dev.playout(8, seed=3, start_states=fin, want_final=True, want_trace=True)
    assert torch.equal(again["final_states"], fin) and (again["trace"].cpu().numpy() == 0xFFFF).all()


def test_playout_replay_sweep():
    """tools/playout_sweep.py: every move and the final scores of 12,000 uniform and 1,200 heuristic whole games equal
    the oracle's replay (profiles/r01w_playout_replay_sweep.json holds a 200,000 + 30,000 game run)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "playout_sweep", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "playout_sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    u = mod.sweep(12000, seed=7, policy=0, first_game=123456)
    h = mod.sweep(1200, seed=7, policy=1, first_game=654321)
    assert u["mismatches"] == 0 and h["mismatches"] == 0
    assert h["mean_score"] > u["mean_score"]


def test_policy_distributions_are_the_references(golden_dir):
    """Not only deterministic parity with the oracle: the SAMPLED distribution of the device policies is the
    reference's.  One mid-game state, 400,000 independent draws of its next move (one per game index): the uniform
    policy hits every legal move equally often, the heuristic policy follows softmax(T=1) of the reference's move
    scores (agents/heuristic_agent.py:222-244; scores are the bit-exact ones of bk_heuristic_scores)."""
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    states = z["states"].view(STATE_DTYPE).reshape(-1)
    st = next(s for s in states if 12 <= int(s["move_count"]) <= 20)
    n = 400_000
    start = dev.to_device(np.repeat(np.atleast_1d(st), n))
    t0 = int(st["move_count"])                                  # any first_turn works: it only names the Philox word
    counts, offsets, ids = dev.movegen(start[:1].contiguous())
    legal = ids.cpu().numpy().astype(np.int64)
    m = len(legal)
    assert m > 50
    scores = dev.heuristic_scores(start[:1].contiguous(), offsets, ids).cpu().numpy()
    p_heur = np.exp(scores - scores.max()); p_heur /= p_heur.sum()
    for policy, p in ((dev.POLICY_UNIFORM, np.full(m, 1.0 / m)), (dev.POLICY_HEURISTIC, p_heur)):
        r = dev.playout(n, seed=123, start_states=start, policy=policy, first_turn=t0, max_turns=t0 + 1, want_trace=True)
        picked = r["trace"][:, t0].cpu().numpy().astype(np.int64)
        pos = np.searchsorted(legal, picked)
        assert (legal[pos] == picked).all()                     # every draw is a legal move
        freq = np.bincount(pos, minlength=m).astype(np.float64)
        expect = n * p
        big = expect >= 20                                      # chi-square over the well-populated cells
        chi2 = float((((freq - expect) ** 2) / expect)[big].sum())
        dof = int(big.sum()) - 1
        assert abs(chi2 - dof) < 6.0 * np.sqrt(2.0 * dof), (policy, chi2, dof)
        assert abs(freq[~big].sum() - expect[~big].sum()) < 6.0 * np.sqrt(expect[~big].sum() + 1.0)


def test_csr_one_pass_equals_two_pass(golden_dir):
    """device.movegen: the one-pass path (fixed rows + compaction) and the two-pass path (count, scan, emit) return the
    same CSR, also when rows overflow the capacity and the wrapper falls back."""
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    st = dev.to_device(z["states"])
    one, two = dev.movegen(st, one_pass=True), dev.movegen(st, one_pass=False)
    assert all(torch.equal(a, b) for a, b in zip(one, two))
    keep = dev.ONE_PASS_CAP
    try:
        dev.ONE_PASS_CAP = 64                                   # most rows overflow: fallback
        small = dev.movegen(st, one_pass=True)
    finally:
        dev.ONE_PASS_CAP = keep
    assert all(torch.equal(a, b) for a, b in zip(small, two))
    empty = dev.movegen(st[:0].contiguous())
    assert empty[2].numel() == 0 and empty[1].tolist() == [0]

#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED Python reference.

Run in the build container only (the reference lives at /root/reference, read-only, and is not
available on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/generate_golden.py

Everything written here is an *output of the reference* (engine/, agents/, mcts/): tables, legal
move sets, replayed games, scores, Zobrist keys, heuristic scores.  The only non-reference
ingredient is the Philox4x32-10 index stream (numpy implementation below) that replaces
RandomAgent's MT19937 draw, so CPU oracle and GPU kernels can replay the very same games.
"""

from __future__ import annotations

import hashlib
import json
import os
import random
import sys

sys.dont_write_bytecode = True
REF = os.environ.get("BLOKUS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

from engine.board import Board, Player  # noqa: E402
from engine.game import BlokusGame  # noqa: E402
from engine.move_generator import LegalMoveGenerator, Move  # noqa: E402
from engine.pieces import ALL_PIECE_ORIENTATIONS  # noqa: E402
from agents.heuristic_agent import HeuristicAgent  # noqa: E402
from mcts.zobrist import ZobristHash  # noqa: E402
from tests.utils_game_states import generate_random_valid_state  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GEN = LegalMoveGenerator()

STATE_DTYPE = np.dtype(
    [("bits", "<u8", (4, 7)), ("used", "<u4", (4,)), ("to_move", "u1"), ("first_mask", "u1"),
     ("move_count", "<u2"), ("reserved", "u1", (12,))]
)

FIRST = {}
_g = 0
for _pid in range(1, 22):
    FIRST[_pid] = _g
    _g += len(ALL_PIECE_ORIENTATIONS[_pid])


def move_id(m: Move) -> int:
    return (FIRST[m.piece_id] + m.orientation) * 400 + m.anchor_row * 20 + m.anchor_col


def id_to_move(i: int) -> Move:
    g, cell = divmod(int(i), 400)
    pid = max(p for p in FIRST if FIRST[p] <= g)
    return Move(pid, g - FIRST[pid], cell // 20, cell % 20)


def pack(board: Board) -> np.ndarray:
    s = np.zeros((), dtype=STATE_DTYPE)
    for p in range(4):
        pl = Player(p + 1)
        raw = int(board.player_bits[pl]).to_bytes(56, "little")
        s["bits"][p] = np.frombuffer(raw, dtype="<u8")
        s["used"][p] = sum(1 << (pid - 1) for pid in board.player_pieces_used[pl])
        if board.player_first_move[pl]:
            s["first_mask"] |= 1 << p
    s["to_move"] = board.current_player.value - 1
    s["move_count"] = board.move_count
    return s


def sorted_ids(board: Board, player: Player, naive: bool = False) -> np.ndarray:
    moves = GEN._get_legal_moves_naive(board, player) if naive else GEN.get_legal_moves(board, player)
    return np.array(sorted(move_id(m) for m in moves), dtype=np.uint16)


def philox4x32_10(ctr, key):
    c = [np.uint64(x) for x in ctr]
    k = [int(key[0]), int(key[1])]
    M0, M1 = 0xD2511F53, 0xCD9E8D57
    c = [int(x) for x in c]
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def draw(seed: int, g: int, t: int, stream: int = 0) -> int:
    """Draw number t of (seed, g, stream): word t%4 of the Philox block with counter (g_lo, g_hi, t//4, stream)."""
    return philox4x32_10([g & 0xFFFFFFFF, g >> 32, t >> 2, stream], [seed & 0xFFFFFFFF, seed >> 32])[t & 3]


# ---------------------------------------------------------------------------------------------
def gen_tables():
    rows = []
    for pid in range(1, 22):
        for o in ALL_PIECE_ORIENTATIONS[pid]:
            rows.append({"piece_id": pid, "index": o.orientation_id, "cells": [list(c) for c in o.offsets]})
            # the generator's own cache must agree (engine/move_generator.py:119-128)
            assert [list(c) for c in GEN.piece_position_cache[pid][o.orientation_id]] == rows[-1]["cells"]
    assert len(rows) == 91
    with open(os.path.join(HERE, "tables.json"), "w") as f:
        json.dump({"source": "engine/pieces.py ALL_PIECE_ORIENTATIONS", "orientations": rows}, f)


def describe_state(board: Board) -> dict:
    d = {
        "grid": board.grid.astype(int).tolist(),
        "to_move": board.current_player.value - 1,
        "move_count": int(board.move_count),
        "used": [sorted(board.player_pieces_used[Player(p + 1)]) for p in range(4)],
        "first_move": [bool(board.player_first_move[Player(p + 1)]) for p in range(4)],
        "frontier": [sorted([list(x) for x in board.get_frontier(Player(p + 1))]) for p in range(4)],
        "has_legal_moves": [bool(GEN.has_legal_moves(board, Player(p + 1))) for p in range(4)],
        "board_score": [int(board.get_score(Player(p + 1))) for p in range(4)],
    }
    g = BlokusGame(enable_telemetry=False)
    g.board = board
    d["game_score"] = [int(g.get_score(Player(p + 1))) for p in range(4)]
    return d


def gen_kat1():
    """benchmarks/benchmark_move_generation.py:163 states: generate_random_valid_state(n, seed=n)."""
    out = []
    for n in (5, 10, 20, 30, 40):
        board, _ = generate_random_valid_state(n, seed=n)
        d = describe_state(board)
        d["n"] = n
        d["legal"] = []
        for p in range(4):
            ids = sorted_ids(board, Player(p + 1))
            naive = sorted_ids(board, Player(p + 1), naive=True)
            assert np.array_equal(ids, naive)
            d["legal"].append(ids.tolist())
        d["packed_hex"] = pack(board).tobytes().hex()
        out.append(d)
        print("kat1", n, [len(x) for x in d["legal"]])
    with open(os.path.join(HERE, "kat1_states.json"), "w") as f:
        json.dump(out, f)


def gen_kat2():
    """tests/fixtures/advanced_metrics_snapshot.json: the reference's recorded 20-move game."""
    with open(os.path.join(REF, "tests/fixtures/advanced_metrics_snapshot.json")) as f:
        snap = json.load(f)
    hist = snap["history"]
    game = BlokusGame(enable_telemetry=False)
    moves = []
    for h in hist:
        a = h["action"] if "action" in h else h
        pl = Player[h["player"]]
        m = Move(a["piece_id"], a["orientation"], a["anchor_row"], a["anchor_col"])
        assert game.board.current_player == pl
        assert game.make_move(m, pl)
        moves.append({"player": pl.value - 1, "move_id": move_id(m),
                      "game_score": [int(game.get_score(Player(p + 1))) for p in range(4)]})
    d = describe_state(game.board)
    d["moves"] = moves
    d["legal"] = [sorted_ids(game.board, Player(p + 1)).tolist() for p in range(4)]
    d["packed_hex"] = pack(game.board).tobytes().hex()
    with open(os.path.join(HERE, "kat2_game.json"), "w") as f:
        json.dump(d, f)
    print("kat2 scores", d["game_score"])


def random_state(num_moves: int, seed: int):
    """Random self-play like tests/utils_game_states.py:12-57 but with the default generator
    (fast) and passes advanced the way callers do (board._update_current_player)."""
    rng = random.Random(seed)
    board = Board()
    made = 0
    stall = 0
    while made < num_moves and stall < 4:
        pl = board.current_player
        moves = GEN.get_legal_moves(board, pl)
        if not moves:
            board._update_current_player()
            stall += 1
            continue
        stall = 0
        moves.sort(key=move_id)
        m = rng.choice(moves)
        assert board.place_piece(m.get_positions(GEN.piece_orientations_cache[m.piece_id]), pl, m.piece_id)
        made += 1
    return board


def gen_random_states(count=240):
    states = np.zeros(count, dtype=STATE_DTYPE)
    offs = [0]
    ids_all = []
    grids = np.zeros((count, 20, 20), dtype=np.int8)
    has = np.zeros((count, 4), dtype=np.uint8)
    bscore = np.zeros((count, 4), dtype=np.int32)
    gscore = np.zeros((count, 4), dtype=np.int32)
    rng = random.Random(20261018)
    for i in range(count):
        depth = rng.choice([0, 1, 2, 3, 4, 6, 9, 13, 18, 24, 31, 39, 48, 56, 64, 72, 80])
        board = random_state(depth, seed=1000 + i)
        states[i] = pack(board)
        grids[i] = board.grid
        g = BlokusGame(enable_telemetry=False)
        g.board = board
        for p in range(4):
            ids = sorted_ids(board, Player(p + 1))
            if i % 16 == 0:
                assert np.array_equal(ids, sorted_ids(board, Player(p + 1), naive=True))
            ids_all.append(ids)
            offs.append(offs[-1] + len(ids))
            has[i, p] = GEN.has_legal_moves(board, Player(p + 1))
            assert bool(has[i, p]) == (len(ids) > 0)
            bscore[i, p] = board.get_score(Player(p + 1))
            gscore[i, p] = g.get_score(Player(p + 1))
        if i % 40 == 0:
            print("random state", i, depth, [len(x) for x in ids_all[-4:]])
    np.savez_compressed(
        os.path.join(HERE, "random_states.npz"),
        states=states.view(np.uint8).reshape(count, 256),
        grids=grids,
        offsets=np.array(offs, dtype=np.int64),
        move_ids=np.concatenate(ids_all).astype(np.uint16),
        has_moves=has, board_scores=bscore, game_scores=gscore,
    )


def gen_games(count=16, seed=42):
    """Complete games through BlokusGame(enable_telemetry=False) with the loop of
    analytics/tournament/arena_runner.py:653-695 and Philox-drawn indices into the SORTED list."""
    traces = np.full((count, 96), 0xFFFF, dtype=np.uint16)
    choices = np.full((count, 96), 0xFFFF, dtype=np.uint16)
    gsc = np.zeros((count, 4), dtype=np.int32)
    bsc = np.zeros((count, 4), dtype=np.int32)
    turns = np.zeros(count, dtype=np.int32)
    finals = np.zeros(count, dtype=STATE_DTYPE)
    for gi in range(count):
        game = BlokusGame(enable_telemetry=False)
        t = 0
        while not game.is_game_over() and t < 96:
            pl = game.get_current_player()
            legal = game.get_legal_moves(pl)
            if not legal:
                game.board._update_current_player()
                game._check_game_over()
                t += 1
                continue
            legal.sort(key=move_id)
            idx = (draw(seed, gi, t) * len(legal)) >> 32
            assert game.make_move(legal[idx], pl)
            traces[gi, t] = move_id(legal[idx])
            choices[gi, t] = idx
            t += 1
        assert game.is_game_over()
        res = game.get_game_result()
        for p in range(4):
            gsc[gi, p] = res.scores[p + 1]
            bsc[gi, p] = game.board.get_score(Player(p + 1))
        turns[gi] = t
        finals[gi] = pack(game.board)
        print("game", gi, "turns", t, "scores", gsc[gi].tolist())
    np.savez_compressed(
        os.path.join(HERE, "games.npz"),
        seed=np.array([seed], dtype=np.uint64), traces=traces, choices=choices,
        game_scores=gsc, board_scores=bsc, turns=turns,
        finals=finals.view(np.uint8).reshape(count, 256),
    )


def gen_zobrist():
    out = {}
    boards = [Board()] + [random_state(d, seed=7000 + d) for d in (1, 5, 17, 33, 60)]
    for seed in (0, 42, 12345):
        z = ZobristHash(seed=seed)
        table = np.concatenate([z.position_player_hashes.reshape(-1), z.player_turn_hashes.reshape(-1),
                                z.piece_used_hashes.reshape(-1)]).astype(np.uint64)
        assert table.size == 2088
        out[str(seed)] = {
            "table_sha256": hashlib.sha256(table.astype("<u8").tobytes()).hexdigest(),
            "first": [int(x) for x in table[:8]],
            "turn": [int(x) for x in z.player_turn_hashes],
            "piece00": int(z.piece_used_hashes[0, 0]),
            "hashes": [int(z.hash_board(b)) for b in boards],
        }
    out["states_hex"] = [pack(b).tobytes().hex() for b in boards]
    with open(os.path.join(HERE, "zobrist.json"), "w") as f:
        json.dump(out, f)
    print("zobrist empty-board hash seed0: %#x" % out["0"]["hashes"][0])


def gen_heuristic():
    agent = HeuristicAgent(seed=0)
    states, offs, ids_all, sc_all = [], [0], [], []
    for d, s in ((0, 1), (3, 2), (8, 3), (21, 4), (29, 5), (30, 6), (31, 7), (45, 8), (62, 9)):
        board = random_state(d, seed=9000 + s)
        pl = board.current_player
        legal = GEN.get_legal_moves(board, pl)
        legal.sort(key=move_id)
        states.append(pack(board))
        ids_all.append(np.array([move_id(m) for m in legal], dtype=np.uint16))
        sc_all.append(np.array([agent._evaluate_move(board, pl, m) for m in legal], dtype=np.float64))
        offs.append(offs[-1] + len(legal))
    np.savez_compressed(
        os.path.join(HERE, "heuristic.npz"),
        states=np.array(states, dtype=STATE_DTYPE).view(np.uint8).reshape(len(states), 256),
        offsets=np.array(offs, dtype=np.int64), move_ids=np.concatenate(ids_all),
        scores=np.concatenate(sc_all),
    )
    print("heuristic", offs[-1], "scored moves")


def gen_mobility():
    """engine/mobility_metrics.py compute_player_mobility_metrics on the random-state corpus (row N4)."""
    from engine.mobility_metrics import compute_player_mobility_metrics

    z = np.load(os.path.join(HERE, "random_states.npz"))
    states = z["states"].view(STATE_DTYPE).reshape(-1)
    n = len(states)
    P = np.zeros((n, 4, 21), dtype=np.int32)
    M = np.zeros((n, 4, 12), dtype=np.float64)
    for i in range(n):
        board = random_state_from_packed(states[i])
        for p in range(4):
            pl = Player(p + 1)
            moves = GEN.get_legal_moves(board, pl)
            for m in moves:
                P[i, p, m.piece_id - 1] += 1
            mm = compute_player_mobility_metrics(moves, sorted(board.player_pieces_used[pl]))
            M[i, p] = [mm.totalPlacements, mm.totalOrientationNormalized, mm.totalCellWeighted, mm.buckets[1],
                       mm.buckets[2], mm.buckets[3], mm.buckets[4], mm.buckets[5], mm.mobilityEntropy,
                       mm.pieceTop1Share, mm.anchorTop1Share, 0.0]
    np.savez_compressed(os.path.join(HERE, "mobility.npz"), piece_counts=P, metrics=M)
    print("mobility", n, "states")


SNAPSHOT_CHECKPOINTS = [0, 8, 16, 24, 32, 40, 48, 56, 64]   # ply 0 + arena_runner.py:44 DEFAULT_SNAPSHOT_PLYS


def _snapshot_matrix(board: Board, turn_index: int, max_turns: int) -> np.ndarray:
    """float64[4, 34]: the feature columns of analytics/winprob/features.py for the four players, computed exactly as
    analytics/tournament/arena_runner.py:522-575 and mcts/learned_evaluator.py:79-95 do."""
    from analytics.winprob.features import (SNAPSHOT_FEATURE_COLUMNS, build_snapshot_runtime_context,
                                            coerce_feature_dict, extract_player_snapshot_features)

    ctx = build_snapshot_runtime_context(board, turn_index=turn_index, max_turns=max_turns)
    out = np.zeros((4, len(SNAPSHOT_FEATURE_COLUMNS)), dtype=np.float64)
    for p in range(4):
        f = coerce_feature_dict(extract_player_snapshot_features(board, player=Player(p + 1), context=ctx, move_generator=GEN))
        out[p] = [f[c] for c in SNAPSHOT_FEATURE_COLUMNS]
    return out


def _frontier_bytes(board: Board) -> np.ndarray:
    fr = np.zeros((4, 400), dtype=np.uint8)
    for p in range(4):
        for (r, c) in board.get_frontier(Player(p + 1)):
            fr[p, r * 20 + c] = 1
    return fr


def gen_snapshots(max_turns=2500):
    """Row N4: snapshot feature rows.  (a) the golden games of games.npz replayed through BlokusGame with the
    arena loop, features captured at the checkpoints exactly where arena_runner.py:625-650,698 captures them
    (incremental frontier sets, turn_index = loop turn count); (b) the random-state corpus evaluated the way
    mcts/learned_evaluator.py does (turn_index = move_count)."""
    from analytics.winprob.features import SNAPSHOT_FEATURE_COLUMNS

    z = np.load(os.path.join(HERE, "games.npz"))
    traces = z["traces"]
    ng, nk, nf = traces.shape[0], len(SNAPSHOT_CHECKPOINTS), len(SNAPSHOT_FEATURE_COLUMNS)
    feats = np.zeros((ng, nk, 4, nf), dtype=np.float64)
    turn_idx = np.full((ng, nk), -1, dtype=np.int32)
    fronts = np.zeros((ng, nk, 4, 400), dtype=np.uint8)
    snaps = np.zeros((ng, nk), dtype=STATE_DTYPE)
    for gi in range(ng):
        game = BlokusGame(enable_telemetry=False)
        hits = set()
        turn_count = 0

        def capture():
            ply = int(game.board.move_count)
            if ply in SNAPSHOT_CHECKPOINTS and ply not in hits:
                k = SNAPSHOT_CHECKPOINTS.index(ply)
                feats[gi, k] = _snapshot_matrix(game.board, turn_count, max_turns)
                turn_idx[gi, k] = turn_count
                fronts[gi, k] = _frontier_bytes(game.board)
                snaps[gi, k] = pack(game.board)
                hits.add(ply)

        capture()
        for t in range(int(z["turns"][gi])):
            pl = game.get_current_player()
            turn_count += 1
            if traces[gi, t] == 0xFFFF:
                game.board._update_current_player()
                game._check_game_over()
                continue
            assert game.make_move(id_to_move(traces[gi, t]), pl)
            capture()
        print("snapshots game", gi, "hit", sorted(hits))
    zs = np.load(os.path.join(HERE, "random_states.npz"))
    states = zs["states"].view(STATE_DTYPE).reshape(-1)
    sfe = np.zeros((len(states), 4, nf), dtype=np.float64)
    for i in range(len(states)):
        b = random_state_from_packed(states[i])
        sfe[i] = _snapshot_matrix(b, int(b.move_count), max_turns)
    np.savez_compressed(os.path.join(HERE, "snapshots.npz"), checkpoints=np.array(SNAPSHOT_CHECKPOINTS, dtype=np.int32),
                        max_turns=np.array([max_turns], dtype=np.int32), features=feats, turn_index=turn_idx,
                        frontiers=np.packbits(fronts, axis=-1, bitorder="little"),
                        snap_states=snaps.view(np.uint8).reshape(ng, nk, 256), state_features=sfe)
    with open(os.path.join(HERE, "snapshot_columns.json"), "w") as f:
        json.dump(SNAPSHOT_FEATURE_COLUMNS, f)
    print("snapshots", ng, "games,", len(states), "states")


def pairwise_input_frame():
    """A small snapshots frame for the pairwise-dataset fixture: reference feature rows of four golden games at three
    checkpoints, made-up final scores (one tie), rows shuffled.  Shared by the generator and tests/test_host_logic.py."""
    import pandas as pd

    z = np.load(os.path.join(HERE, "snapshots.npz"))
    with open(os.path.join(HERE, "snapshot_columns.json")) as f:
        cols = json.load(f)
    rng = np.random.RandomState(0)
    rows = []
    for g in range(4):
        scores = [int(v) for v in rng.randint(40, 90, 4)]
        if g == 1:
            scores[2] = scores[0]
        for k in (1, 3, 5):
            for p in range(4):
                r = {"run_id": "r", "game_id": f"r_g{g:04d}", "checkpoint_index": k, "player_id": p + 1,
                     "agent_name": f"ag{(p + g) % 4}", "final_score": scores[p]}
                r.update({c: float(z["features"][g, k, p, i]) for i, c in enumerate(cols)})
                rows.append(r)
    return pd.DataFrame(rows).sample(frac=1.0, random_state=1), cols


def gen_pairwise():
    """analytics/winprob/dataset.py build_pairwise_dataset / split_pairwise_by_game on pairwise_input_frame()."""
    from analytics.winprob.dataset import build_pairwise_dataset, split_pairwise_by_game

    df, cols = pairwise_input_frame()
    out, meta = build_pairwise_dataset(df)
    train, test, split = split_pairwise_by_game(out, test_size=0.25, seed=3)
    np.savez_compressed(
        os.path.join(HERE, "pairwise.npz"), X=out[cols].to_numpy(dtype=np.float64), label=out["label"].to_numpy(np.int64),
        player_i=out["player_i_id"].to_numpy(np.int64), player_j=out["player_j_id"].to_numpy(np.int64),
        checkpoint=out["checkpoint_index"].to_numpy(np.int64), occupancy=out["phase_board_occupancy"].to_numpy(np.float64),
        game=np.array(out["game_id"].tolist()), agent_i=np.array(out["agent_i"].tolist()), agent_j=np.array(out["agent_j"].tolist()),
        bucket=np.array(out["phase_bucket"].tolist()), columns=np.array(list(out.columns)),
        meta=np.array([json.dumps(meta), json.dumps(split)]), train_index=train.index.to_numpy(np.int64),
        test_index=test.index.to_numpy(np.int64))
    print("pairwise", meta, split)


# ---------------------------------------------------------------------------------------------
# UCT search of the reference itself (mcts/mcts_agent.py), made reproducible without touching its code
# ---------------------------------------------------------------------------------------------
class _AscendingGenerator:
    """Context manager: the SHARED generator (engine/move_generator.py get_shared_generator, used by MCTSNode and
    MCTSAgent) returns its legal moves in ascending canonical id.  The reference's order is CPython set order; any
    order is a valid run of its algorithm, this one is the order the oracle and the kernel use."""

    def __enter__(self):
        from engine.move_generator import get_shared_generator

        self.gen = get_shared_generator()
        orig = self.gen.get_legal_moves            # bound method of the unmodified class

        def ascending(board, player):
            return sorted(orig(board, player), key=move_id)

        self.gen.get_legal_moves = ascending       # instance attribute shadows the method
        return self

    def __exit__(self, *a):
        del self.gen.get_legal_moves


class PhiloxIndexAgent:
    """Rollout agent (the `rollout_agent` argument of MCTSAgent, mcts_agent.py:207,275-278): RandomAgent's uniform
    index choice with the draw taken from Philox stream 1 of (seed, tree): draw number n is this agent's n-th call."""

    def __init__(self, seed: int, tree: int = 0):
        self.seed, self.tree, self.n = seed, tree, 0

    def select_action(self, board, player, legal_moves):
        r = draw(self.seed, self.tree, self.n, 1)
        self.n += 1
        return legal_moves[(r * len(legal_moves)) >> 32]

    def set_seed(self, seed):
        self.seed = seed


def _device_order(ids):
    """Cumulative order of the device's categorical draw (csrc/heuristic.cuh): the moves are grouped into entries (anchor row,
    orientation), entries numbered row by row; lane L of 32 owns entries L, L+32, ... (columns ascending inside an entry);
    the cumulative sum runs over lane 0's moves, then lane 1's, ...  `ids` = ascending move ids; returns indices into it."""
    ids = np.asarray(ids, dtype=np.int64)
    o, row, col = ids // 400, (ids % 400) // 20, ids % 20
    _, entry = np.unique(row * 128 + o, return_inverse=True)         # entry number in (row, orientation) order
    return np.lexsort((col, entry, entry % 32))


class _PhiloxChoice:
    """Stands in for HeuristicAgent.rng (agents/heuristic_agent.py:29,65): choice(n, p=probs) turns the next Philox
    draw of stream 1 into a categorical sample.  The probabilities are the reference's own softmax; the cumulative
    order is the device's (_device_order; `ids` = the ascending move ids of the list the probabilities belong to, set by
    the caller before select_action runs)."""

    stream = 1

    def __init__(self, seed: int, tree: int = 0):
        self.seed, self.tree, self.n, self.ids = seed, tree, 0, None

    def choice(self, n, p):
        r = draw(self.seed, self.tree, self.n, self.stream)
        self.n += 1
        assert self.ids is not None and len(self.ids) == n
        order = _device_order(self.ids)
        cdf = np.cumsum(np.asarray(p, dtype=np.float64)[order])
        target = cdf[-1] * (r * 2.0 ** -32)
        k = int(np.searchsorted(cdf, target, side="right"))
        return int(order[min(k, n - 1)])


def _root_children(root):
    ch = sorted(root.children, key=lambda c: move_id(c.move))
    return (np.array([move_id(c.move) for c in ch], dtype=np.uint16), np.array([c.visits for c in ch], dtype=np.int32),
            np.array([float(c.total_reward) for c in ch], dtype=np.float64))


def _tree_digest(root):
    """(nodes, max depth, sum over nodes of visits*(depth+1)) -- a cheap fingerprint of the whole tree's shape."""
    n = d_max = acc = 0
    stack = [(root, 0)]
    while stack:
        nd, d = stack.pop()
        n += 1
        d_max = max(d_max, d)
        acc += nd.visits * (d + 1)
        stack.extend((c, d + 1) for c in nd.children)
    return np.array([n, d_max, acc], dtype=np.int64)


def _fit_logreg_artifact(path):
    """A fitted pairwise_logreg artifact in the reference's format (analytics/winprob/train.py writes
    {"model_type", "feature_columns", "pipeline"}): StandardScaler + LogisticRegression on pairwise.npz."""
    import joblib
    from sklearn.linear_model import LogisticRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    from analytics.winprob.features import SNAPSHOT_FEATURE_COLUMNS

    z = np.load(os.path.join(HERE, "pairwise.npz"))
    pipe = Pipeline([("scaler", StandardScaler()), ("clf", LogisticRegression(max_iter=500))]).fit(z["X"], z["label"])
    joblib.dump({"model_type": "pairwise_logreg", "feature_columns": list(SNAPSHOT_FEATURE_COLUMNS), "pipeline": pipe}, path)
    sc, clf = pipe.named_steps["scaler"], pipe.named_steps["clf"]
    w = clf.coef_[0] / sc.scale_
    return np.concatenate([w, [clf.intercept_[0] - float(np.dot(w, sc.mean_))]]).astype(np.float64)


def gen_mcts(iterations=200, seed=42):
    """Runs the UNMODIFIED MCTSAgent / MCTSNode (mcts/mcts_agent.py:19-191, 304-582) and records the root statistics.
    Only two inputs are pinned from outside, through the reference's own extension points: the order of the shared
    generator's move list (ascending id) and the rollout agent's random draw (Philox).  Selection, UCB1, pop-last
    expansion, transposition-table probe/store, rollout loop, reward and back-propagation are the reference's code."""
    import tempfile

    from mcts.mcts_agent import MCTSAgent, MCTSNode

    roots = [("empty", Board())] + [("kat1_%d" % n, generate_random_valid_state(n, seed=n)[0]) for n in (5, 10, 20, 30, 40)]
    out = {"iterations": np.array([iterations]), "seed": np.array([seed], dtype=np.uint64)}
    names = []

    def record(tag, board, player, agent, its, frontier=True):
        root = MCTSNode(board, player)
        agent.iterations = its
        agent._run_mcts_with_iterations(root)
        mv, vis, rew = _root_children(root)
        out[tag + "/state"] = np.frombuffer(_packed_for(board, player).tobytes(), dtype=np.uint8)
        out[tag + "/frontier"] = _frontier_bytes(board)
        out[tag + "/moves"], out[tag + "/visits"], out[tag + "/reward"] = mv, vis, rew
        out[tag + "/digest"] = _tree_digest(root)
        out[tag + "/tt"] = np.array([agent.stats["transposition_hits"],
                                     len(agent.transposition_table.table) if agent.transposition_table else 0], dtype=np.int64)
        best = root.get_best_move()
        out[tag + "/best"] = np.array([move_id(best)], dtype=np.uint16)
        names.append(tag)
        print(tag, "n_root", len(mv), "visited", int((vis > 0).sum()), "digest", out[tag + "/digest"].tolist(),
              "tt", out[tag + "/tt"].tolist(), flush=True)
        return root, best

    with _AscendingGenerator():
        for name, board in roots:
            player = board.current_player
            # 1. uniform Philox rollouts, no transposition table
            ag = MCTSAgent(iterations=iterations, rollout_agent=PhiloxIndexAgent(seed), use_transposition_table=False, seed=seed)
            record("uct/%s" % name, board, player, ag, iterations)
            # 2. with the transposition table; then a SECOND search of the same agent (same table, mcts_agent.py:337-339
            #    clears it only past 500 000 entries) on the position after the chosen move
            rag = PhiloxIndexAgent(seed)
            ag = MCTSAgent(iterations=iterations, rollout_agent=rag, use_transposition_table=True, seed=seed)
            _, best = record("tt1/%s" % name, board, player, ag, iterations)
            nxt = board.copy()
            nxt.place_piece(ag._get_move_positions(best), player, best.piece_id, validate=False)
            rag.tree, rag.n = 1, 0                  # the agent's second search draws from the next stream block
            record("tt2/%s" % name, nxt, nxt.current_player, ag, iterations)
        # 3. the default rollout policy: HeuristicAgent with its RandomState replaced by the Philox categorical draw
        for name, board in roots[:3]:
            from agents.heuristic_agent import HeuristicAgent as _H

            class _HIds(_H):                       # hands the move ids of the list to the draw; select_action itself is the reference's
                def select_action(self, board, player, legal_moves):
                    self.rng.ids = [move_id(m) for m in legal_moves]
                    return super().select_action(board, player, legal_moves)

            h = _HIds(seed=seed)
            h.rng = _PhiloxChoice(seed)
            ag = MCTSAgent(iterations=60, rollout_agent=h, use_transposition_table=False, seed=seed)
            record("heur/%s" % name, board, board.current_player, ag, 60)
        # 4. learned options: the reference's dummy artifact, and a fitted pairwise_logreg pipeline
        with tempfile.TemporaryDirectory() as td:
            dummy = os.path.join(td, "dummy_model.json")
            with open(dummy, "w") as f:
                f.write("{}")
            art = os.path.join(td, "pairwise_logreg.pkl")
            out["model"] = _fit_logreg_artifact(art)
            for name, board in roots[:3]:
                player = board.current_player
                ag = MCTSAgent(iterations=iterations, rollout_agent=PhiloxIndexAgent(seed), use_transposition_table=False, seed=seed,
                               learned_model_path=dummy, leaf_evaluation_enabled=True)
                record("dummy/%s" % name, board, player, ag, iterations)
                ag = MCTSAgent(iterations=80, rollout_agent=PhiloxIndexAgent(seed), use_transposition_table=False, seed=seed,
                               learned_model_path=art, leaf_evaluation_enabled=True)
                record("leaf/%s" % name, board, player, ag, 80)
                ag = MCTSAgent(iterations=80, rollout_agent=PhiloxIndexAgent(seed), use_transposition_table=True, seed=seed,
                               learned_model_path=art, leaf_evaluation_enabled=True, progressive_bias_enabled=True,
                               progressive_bias_weight=0.25)
                record("bias/%s" % name, board, player, ag, 80)
            for name, board in roots[:2]:
                for mode in ("prob", "logit"):
                    ag = MCTSAgent(iterations=40, rollout_agent=PhiloxIndexAgent(seed), use_transposition_table=False, seed=seed,
                                   learned_model_path=art, potential_shaping_enabled=True, potential_shaping_gamma=0.9,
                                   potential_shaping_weight=2.0, potential_mode=mode, max_rollout_moves=6)
                    record("shape_%s/%s" % (mode, name), board, board.current_player, ag, 40)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "mcts.npz"), **out)
    print("mcts", len(names), "searches")


def gen_mcts_deep(seed=42):
    """More searches of the unmodified MCTSAgent on late-game roots with few legal moves, where 300 iterations build
    trees several plies deep (UCB1 selection below the root, terminal nodes, in-search transpositions).  Added to the
    entries gen_mcts wrote."""
    from agents.heuristic_agent import HeuristicAgent as _H
    from mcts.mcts_agent import MCTSAgent, MCTSNode

    path = os.path.join(HERE, "mcts.npz")
    out = {k: v for k, v in np.load(path).items()}
    names = [str(n) for n in out["names"] if not str(n).split("/")[0].startswith("deep")]
    states = np.load(os.path.join(HERE, "random_states.npz"))["states"].view(STATE_DTYPE).reshape(-1)

    def record(tag, board, player, agent, its):
        root = MCTSNode(board, player)
        agent.iterations = its
        agent._run_mcts_with_iterations(root)
        mv, vis, rew = _root_children(root)
        out[tag + "/state"] = np.frombuffer(_packed_for(board, player).tobytes(), dtype=np.uint8)
        out[tag + "/frontier"] = _frontier_bytes(board)
        out[tag + "/moves"], out[tag + "/visits"], out[tag + "/reward"] = mv, vis, rew
        out[tag + "/digest"] = _tree_digest(root)
        out[tag + "/tt"] = np.array([agent.stats["transposition_hits"],
                                     len(agent.transposition_table.table) if agent.transposition_table else 0], dtype=np.int64)
        out[tag + "/best"] = np.array([move_id(root.get_best_move())], dtype=np.uint16)
        names.append(tag)
        print(tag, "children", len(mv), "digest", out[tag + "/digest"].tolist(), "tt", out[tag + "/tt"].tolist(), flush=True)

    with _AscendingGenerator():
        for idx in (42, 74, 80, 54):
            board = random_state_from_packed(states[idx])
            player = board.current_player
            name = "rs%d" % idx
            ag = MCTSAgent(iterations=300, rollout_agent=PhiloxIndexAgent(seed), use_transposition_table=False, seed=seed)
            record("deep_uct/" + name, board, player, ag, 300)
            ag = MCTSAgent(iterations=300, rollout_agent=PhiloxIndexAgent(seed), use_transposition_table=True, seed=seed)
            record("deep_tt/" + name, board, player, ag, 300)
            class _HIds(_H):                       # hands the move ids of the list to the draw; select_action itself is the reference's
                def select_action(self, board, player, legal_moves):
                    self.rng.ids = [move_id(m) for m in legal_moves]
                    return super().select_action(board, player, legal_moves)

            h = _HIds(seed=seed)
            h.rng = _PhiloxChoice(seed)
            ag = MCTSAgent(iterations=150, rollout_agent=h, use_transposition_table=True, seed=seed)
            record("deep_heur/" + name, board, player, ag, 150)
    out["names"] = np.array(names)
    np.savez_compressed(path, **out)
    print("mcts", len(names), "searches")


def gen_mcts_gbt(seed=42):
    """The evaluator's other artifact type, "pairwise_gbt_phase" (mcts/learned_evaluator.py:112-127), through the
    unmodified MCTSAgent with leaf evaluation: an "early" phase model plus a fallback model (used for the mid-game
    root), scikit-learn GradientBoostingClassifier fitted on pairwise.npz.  The artifact is stored flattened
    (reinforcementlearning_blokus_b200.winprob.flatten_gbt_phase_models -- checked against scikit-learn on its own in
    tests/test_host_logic.py) next to the searches; added to the entries gen_mcts wrote."""
    import tempfile

    import joblib
    from sklearn.ensemble import GradientBoostingClassifier

    from analytics.winprob.features import SNAPSHOT_FEATURE_COLUMNS
    from mcts.mcts_agent import MCTSAgent, MCTSNode

    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from reinforcementlearning_blokus_b200.winprob import flatten_gbt_phase_models

    path = os.path.join(HERE, "mcts.npz")
    out = {k: v for k, v in np.load(path).items()}
    names = [str(n) for n in out["names"] if not str(n).startswith("gbt/")]
    z = np.load(os.path.join(HERE, "pairwise.npz"))
    early = GradientBoostingClassifier(n_estimators=24, max_depth=3, random_state=0).fit(z["X"], z["label"])
    fallback = GradientBoostingClassifier(n_estimators=40, max_depth=2, learning_rate=0.2, random_state=1).fit(z["X"][::2], z["label"][::2])
    artifact = {"model_type": "pairwise_gbt_phase", "feature_columns": list(SNAPSHOT_FEATURE_COLUMNS),
                "phase_models": {"early": early}, "fallback_model": fallback}
    out["gbt_model"] = flatten_gbt_phase_models(artifact)
    roots = [("empty", Board()), ("kat1_40", generate_random_valid_state(40, seed=40)[0])]
    with tempfile.TemporaryDirectory() as td, _AscendingGenerator():
        art = os.path.join(td, "gbt.pkl")
        joblib.dump(artifact, art)
        for name, board in roots:
            player = board.current_player
            ag = MCTSAgent(iterations=80, rollout_agent=PhiloxIndexAgent(seed), use_transposition_table=False, seed=seed,
                           learned_model_path=art, leaf_evaluation_enabled=True)
            root = MCTSNode(board, player)
            ag._run_mcts_with_iterations(root)
            tag = "gbt/" + name
            mv, vis, rew = _root_children(root)
            out[tag + "/state"] = np.frombuffer(_packed_for(board, player).tobytes(), dtype=np.uint8)
            out[tag + "/frontier"] = _frontier_bytes(board)
            out[tag + "/moves"], out[tag + "/visits"], out[tag + "/reward"] = mv, vis, rew
            out[tag + "/digest"] = _tree_digest(root)
            out[tag + "/tt"] = np.array([0, 0], dtype=np.int64)
            out[tag + "/best"] = np.array([move_id(root.get_best_move())], dtype=np.uint16)
            names.append(tag)
            print(tag, "children", len(mv), "digest", out[tag + "/digest"].tolist(), "reward sum", float(rew.sum()), flush=True)
    out["names"] = np.array(names)
    np.savez_compressed(path, **out)
    print("mcts", len(names), "searches")


def gen_heuristic_games(count=6, seed=42):
    """Complete games where EVERY move is chosen by the unmodified HeuristicAgent.select_action (agents/heuristic_agent.py:
    39-66: its scores, its softmax), with only its RandomState replaced by the Philox categorical draw of turn t
    (stream 0 of (seed, game), the playout kernel's stream) -- the arena loop of analytics/tournament/arena_runner.py:653-695."""
    from agents.heuristic_agent import HeuristicAgent as _H

    traces = np.full((count, 96), 0xFFFF, dtype=np.uint16)
    gsc = np.zeros((count, 4), dtype=np.int32)
    turns = np.zeros(count, dtype=np.int32)

    class TurnChoice(_PhiloxChoice):
        stream = 0                                             # stream 0, draw number = turn (set by the loop below)

    for gi in range(count):
        game = BlokusGame(enable_telemetry=False)
        agent = _H(seed=seed)
        agent.rng = TurnChoice(seed, gi)
        t = 0
        while not game.is_game_over() and t < 96:
            pl = game.get_current_player()
            legal = game.get_legal_moves(pl)
            if not legal:
                game.board._update_current_player()
                game._check_game_over()
                t += 1
                continue
            legal.sort(key=move_id)
            agent.rng.n = t
            agent.rng.ids = [move_id(m) for m in legal]
            mv = agent.select_action(game.board, pl, legal)
            assert game.make_move(mv, pl)
            traces[gi, t] = move_id(mv)
            t += 1
        res = game.get_game_result()
        gsc[gi] = [res.scores[p + 1] for p in range(4)]
        turns[gi] = t
        print("heuristic game", gi, "turns", t, "scores", gsc[gi].tolist(), flush=True)
    np.savez_compressed(os.path.join(HERE, "heuristic_games.npz"), seed=np.array([seed], dtype=np.uint64), traces=traces,
                        game_scores=gsc, turns=turns)


def gen_tt_eviction():
    """mcts/zobrist.py TranspositionTable with a small max_size: a sequence of puts (one of them overwrites an old key)
    and what is left afterwards -- pins the evict-oldest-tenth rule (:196-200) and dict semantics of put."""
    from mcts.zobrist import TranspositionTable

    rng = np.random.default_rng(0)
    keys = [int(k) for k in rng.integers(1, 2**63, size=160)]
    tt = TranspositionTable(max_size=50)
    ops = []
    for i, k in enumerate(keys):
        tt.put(k, {"reward": float(i)})
        ops.append([k, float(i)])
        if i in (30, 77, 120):
            tt.put(keys[i - 28], {"reward": -float(i)})          # overwrite an entry that may or may not still be there
            ops.append([keys[i - 28], -float(i)])
    left = {str(k): v["reward"] for k, v in tt.table.items()}
    with open(os.path.join(HERE, "tt_eviction.json"), "w") as f:
        json.dump({"max_size": 50, "ops": ops, "left": left, "order": [str(k) for k in tt.table]}, f)
    print("tt_eviction", len(ops), "puts,", len(left), "entries left")


def _packed_for(board: Board, player: Player) -> np.ndarray:
    s = pack(board)
    s["to_move"] = player.value - 1
    return s


def gen_telemetry():
    """engine/telemetry.py collect_all_player_metrics (fast and full sampling) on boards replayed from the golden games,
    i.e. with the reference's incremental frontier sets.  The generator handed to it lists moves in ascending id (the
    order simulate_mobility_stability's seeded shuffle then starts from)."""
    from engine.telemetry import collect_all_player_metrics

    class Ascending:
        def __init__(self):
            self.piece_orientations_cache = GEN.piece_orientations_cache

        def get_legal_moves(self, board, player):
            return sorted(GEN.get_legal_moves(board, player), key=move_id)

    z = np.load(os.path.join(HERE, "games.npz"))
    traces = z["traces"]
    rows = []
    for gi, plies in ((0, (0, 3, 9, 22, 41, 60)), (1, (5, 14, 33, 52)), (2, (1, 27, 47, 70)), (5, (12, 36, 64))):
        game = BlokusGame(enable_telemetry=False)
        t = 0
        for stop in plies:
            while t < stop and t < int(z["turns"][gi]):
                pl = game.get_current_player()
                if traces[gi, t] == 0xFFFF:
                    game.board._update_current_player()
                else:
                    assert game.make_move(id_to_move(traces[gi, t]), pl)
                t += 1
            board = game.board
            for fast in (True, False):
                if not fast and stop % 2:
                    continue
                m = collect_all_player_metrics(board, Ascending(), fast_mode=fast)
                rows.append({"game": gi, "turn": t, "fast_mode": fast, "packed_hex": pack(board).tobytes().hex(),
                             "frontier": [sorted([list(x) for x in board.get_frontier(Player(p + 1))]) for p in range(4)],
                             "metrics": m, "no_generator": collect_all_player_metrics(board, None) if fast else None})
            print("telemetry game", gi, "turn", t, flush=True)
    # BlokusGame(enable_telemetry=True): the per-move payload of engine/game.py:72-141 for the first plies of golden game 3
    moves = []
    with _AscendingGenerator():
        game = BlokusGame(enable_telemetry=True, telemetry_fast_mode=True)
        for t in range(7):
            pl = game.get_current_player()
            assert game.make_move(id_to_move(traces[3, t]), pl)
            tel = game.game_history[-1]["telemetry"]
            moves.append({"move_id": int(traces[3, t]), "telemetry": {k: tel[k] for k in
                          ("ply", "moverId", "moveId", "deltaSelf", "deltaOppTotal", "deltaOppByPlayer", "deltaAdvantage", "winProxyDelta")}})
    with open(os.path.join(HERE, "telemetry.json"), "w") as f:
        json.dump({"source": "engine/telemetry.py collect_all_player_metrics", "rows": rows, "game_moves": moves}, f)
    print("telemetry", len(rows), "rows")


def random_state_from_packed(s) -> Board:
    """Rebuild a reference Board from a packed state (the recipe of engine-service/app.py:27-50)."""
    board = Board()
    for p in range(4):
        pl = Player(p + 1)
        m = int.from_bytes(np.ascontiguousarray(s["bits"][p], dtype="<u8").tobytes(), "little")
        board.player_bits[pl] = m
        board.occupied_bits |= m
        for i in range(400):
            if (m >> i) & 1:
                board.grid[i // 20, i % 20] = pl.value
        used = int(s["used"][p])
        board.player_pieces_used[pl] = {pid for pid in range(1, 22) if (used >> (pid - 1)) & 1}
        board.player_first_move[pl] = bool((int(s["first_mask"]) >> p) & 1)
    for pl in Player:
        if board.player_pieces_used[pl]:
            board.player_frontiers[pl] = board._compute_full_frontier(pl)
        else:
            c = board.player_start_corners[pl]
            board.player_frontiers[pl] = {(c.row, c.col)}
    board.current_player = Player(int(s["to_move"]) + 1)
    board.move_count = int(s["move_count"])
    return board


if __name__ == "__main__":
    which = sys.argv[1:] or ["tables", "kat1", "kat2", "random", "games", "zobrist", "heuristic", "mobility", "snapshots", "pairwise", "mcts", "mcts_deep", "mcts_gbt", "heuristic_games", "tt_eviction", "telemetry"]
    for w in which:
        {"tables": gen_tables, "kat1": gen_kat1, "kat2": gen_kat2, "random": gen_random_states,
         "games": gen_games, "zobrist": gen_zobrist, "heuristic": gen_heuristic, "mobility": gen_mobility,
         "snapshots": gen_snapshots, "pairwise": gen_pairwise, "mcts": gen_mcts, "mcts_deep": gen_mcts_deep, "mcts_gbt": gen_mcts_gbt, "heuristic_games": gen_heuristic_games, "tt_eviction": gen_tt_eviction, "telemetry": gen_telemetry}[w]()

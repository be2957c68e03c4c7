"""Row N3: engine/telemetry.py (simulate_mobility_stability, compute_player_metrics, collect_all_player_metrics).

Fixture: tests/golden/telemetry.json, written by the unmodified reference on boards replayed from the golden games
(incremental frontier sets) with its generator listing moves in ascending id.
CPU: the host-side parts (geometric metrics, sampling plan, statistics) -- legal lists and after-move counts come from
the C oracle there.  GPU: the product path end to end (four launches per board)."""

import json
import math
import os

import numpy as np
import pytest

from oracle import orc
from reinforcementlearning_blokus_b200 import pieces as P
from reinforcementlearning_blokus_b200 import telemetry as T
from reinforcementlearning_blokus_b200.board import Board, Player
from reinforcementlearning_blokus_b200.state import STATE_DTYPE

# effectiveFrontier is a float sum over a Python SET in the reference (order = CPython hash-table history); every
# other entry is order-free and must be bit-identical
ORDER_DEPENDENT = {"effectiveFrontier": 1e-12}


def _rows(golden_dir):
    with open(os.path.join(golden_dir, "telemetry.json")) as f:
        return json.load(f)["rows"]


def _board(row) -> Board:
    b = Board.from_packed(np.frombuffer(bytes.fromhex(row["packed_hex"]), dtype=STATE_DTYPE)[0])
    for k, p in enumerate(Player):                      # the reference's incremental frontier sets
        b.player_frontier_bits[p] = sum(1 << (r * 20 + c) for r, c in row["frontier"][k])
    return b


def _check(got, want, where):
    assert set(got) == set(want), (where, set(got) ^ set(want))
    for k, w in want.items():
        if k in ORDER_DEPENDENT:
            assert math.isclose(got[k], w, rel_tol=ORDER_DEPENDENT[k], abs_tol=1e-12), (where, k, got[k], w)
        else:
            assert got[k] == w, (where, k, got[k], w)


def test_static_and_dead_space_metrics_match_reference(golden_dir):
    for row in _rows(golden_dir):
        if row["no_generator"] is None:
            continue
        b = _board(row)
        got = T.collect_all_player_metrics(b, None)              # no generator: no device work at all
        for p in Player:
            _check(got[p.name], row["no_generator"][p.name], (row["game"], row["turn"], p.name))


def test_sampling_and_statistics_match_reference_with_oracle_counts(golden_dir):
    """simulate_mobility_stability + the generator-dependent entries, with the C oracle standing in for the device."""
    for row in _rows(golden_dir):
        b = _board(row)
        ob = orc.OracleBoard.from_packed(b.to_packed())
        lists = [ob.legal_moves(q) for q in range(4)]
        k = T.TelemetryConfig.MOBILITY_SAMPLES_K if row["fast_mode"] else 8
        for p in Player:
            q = p.value - 1
            after = []
            for mover, mid in (T._sample_plan(lists, q, k) if len(lists[q]) else []):
                c = ob.copy()
                assert c.place(mover, mid)
                after.append(len(c.legal_moves(q)))
            stab = T._stability_from_samples(len(lists[q]), after)
            pc = np.zeros(21, dtype=np.int64)
            anchors = {}
            for mid in lists[q]:
                pc[P.ORIENTATIONS[int(mid) // 400].piece_id - 1] += 1
                anchors[int(mid) % 400] = anchors.get(int(mid) % 400, 0) + 1
            got = T._static_block(b, p)
            got.update(T._movegen_block(b, p, pc, max(anchors.values(), default=0), stab))
            want = {k2: v for k2, v in row["metrics"][p.name].items() if not k2.startswith("deadSpace")}
            _check(got, want, (row["game"], row["turn"], row["fast_mode"], p.name))


@pytest.mark.gpu
def test_device_telemetry_matches_reference(golden_dir):
    pytest.importorskip("torch")
    from reinforcementlearning_blokus_b200 import get_shared_generator
    from reinforcementlearning_blokus_b200 import device as dev

    gen = get_shared_generator()
    for row in _rows(golden_dir):
        b = _board(row)
        n0 = dev.launches()
        got = T.collect_all_player_metrics(b, gen, fast_mode=row["fast_mode"])
        assert dev.launches() - n0 <= 10                  # generation (fixed rows, scan, pack), piece counts, step, count
        for p in Player:
            _check(got[p.name], row["metrics"][p.name], (row["game"], row["turn"], row["fast_mode"], p.name))
        one = T.compute_player_metrics(b, b.current_player, gen, fast_mode=row["fast_mode"])
        want = {k: v for k, v in row["metrics"][b.current_player.name].items() if not k.startswith("deadSpace")}
        _check(one, want, "single player")
        stab = T.simulate_mobility_stability(b, b.current_player, gen, row["fast_mode"])
        for k, v in stab.items():
            assert v == row["metrics"][b.current_player.name][k]


@pytest.mark.gpu
def test_game_records_the_references_telemetry_payload(golden_dir):
    """BlokusGame(enable_telemetry=True): deltas of engine/game.py:72-141 / engine/telemetry.py:200-276 per move."""
    pytest.importorskip("torch")
    from reinforcementlearning_blokus_b200 import BlokusGame, Move

    with open(os.path.join(golden_dir, "telemetry.json")) as f:
        moves = json.load(f)["game_moves"]
    game = BlokusGame(enable_telemetry=True, telemetry_fast_mode=True)
    for mv in moves:
        assert game.make_move(Move.from_id(mv["move_id"]), game.get_current_player())
        got, want = game.game_history[-1]["telemetry"], mv["telemetry"]
        assert (got["ply"], got["moverId"], got["moveId"]) == (want["ply"], want["moverId"], want["moveId"])
        for block in ("deltaSelf", "deltaOppTotal", "deltaAdvantage"):
            assert set(got[block]) == set(want[block])
            for k, w in want[block].items():
                assert math.isclose(got[block][k], w, rel_tol=1e-9, abs_tol=1e-9), (block, k)
        for name, d in want["deltaOppByPlayer"].items():
            for k, w in d.items():
                assert math.isclose(got["deltaOppByPlayer"][name][k], w, rel_tol=1e-9, abs_tol=1e-9)
        assert math.isclose(got["winProxyDelta"], want["winProxyDelta"], rel_tol=1e-9, abs_tol=1e-9)
        assert len(got["before"]) == len(got["after"]) == 4
    assert "telemetry" not in BlokusGame().__dict__

"""Arena: the caller on top of the hot path.

Two ways to play match-ups, both producing game records with the keys of the reference's ``games.jsonl``
(analytics/tournament/arena_runner.py:752-777):

* ``run_single_game`` -- the reference's turn loop (arena_runner.py:653-700) over agent OBJECTS of this package
  (any mix of RandomAgent / HeuristicAgent / FastMCTSAgent / MCTSAgent): legal moves, pass handling, invalid-move
  handling, max_turns truncation, per-agent move statistics.
* ``run_device_arena`` -- whole match-ups between the two built-in playout policies as ONE device batch: every
  game gets its seat assignment (round-robin or seeded shuffle, arena_runner.py:279-292) and game seed
  (``stable_hash_int``, :241-255) exactly as the reference derives them, and the games are played by
  ``bk_playout_seats`` with one policy per seat.

Seeds and seat assignments follow the reference bit for bit; the moves of the built-in policies are drawn
with Philox (DESIGN.md "Random streams"), so individual games differ from a reference run while the rules,
scoring and bookkeeping are identical.
"""

from __future__ import annotations

import hashlib
import json
import random
import time
import traceback
from typing import Any, Dict, Mapping, Optional, Sequence, Tuple

import numpy as np

from .board import Player
from .game import BlokusGame

POLICY_OF = {"random": 0, "heuristic": 1}


# ---- deterministic seeds / seats (arena_runner.py:241-292) -------------------------------------------
def stable_hash_int(*parts: Any, mod: int = 2**31 - 1) -> int:
    payload = "|".join(str(part) for part in parts).encode("utf-8")
    return int(hashlib.sha256(payload).hexdigest()[:16], 16) % mod


def game_seed_from_run_seed(run_seed: int, game_index: int) -> int:
    return stable_hash_int(run_seed, game_index, "game_seed")


def agent_seed(run_seed: int, game_index: int, agent_name: str) -> int:
    return stable_hash_int(run_seed, game_index, agent_name, "agent_seed")


def seat_assignment_for_game(agent_names: Sequence[str], game_index: int, game_seed: int, seat_policy: str) -> Dict[str, str]:
    ordered = list(agent_names)
    if seat_policy == "round_robin":
        shift = game_index % len(ordered)
        ordered = ordered[shift:] + ordered[:shift]
    else:
        random.Random(stable_hash_int(game_seed, "seat_assignment")).shuffle(ordered)
    return {str(player.value): ordered[idx] for idx, player in enumerate(Player)}


def compute_ranks(scores_by_player: Mapping[str, int]) -> Dict[str, int]:
    """Dense ranking as in the reference: 1 + number of DISTINCT better scores (ties share a rank, no gaps)."""
    distinct = sorted(set(scores_by_player.values()), reverse=True)
    return {pid: 1 + distinct.index(s) for pid, s in scores_by_player.items()}


def _record(run_id, game_index, game_seed, seat_assignment, seat_policy, scores_by_player, moves_made, turn_count,
            passes, invalid_actions, duration, truncated, per_agent_stats, error) -> Dict[str, Any]:
    top = max(scores_by_player.values())
    winner_ids = [int(pid) for pid, s in scores_by_player.items() if s == top]
    ranks = compute_ranks(scores_by_player)
    return {
        "run_id": run_id, "game_id": f"{run_id}_g{game_index:04d}", "game_index": game_index, "game_seed": game_seed,
        "seat_assignment": dict(seat_assignment), "seat_policy": seat_policy, "winner_ids": winner_ids,
        "winner_agents": [seat_assignment[str(p)] for p in winner_ids], "is_tie": len(winner_ids) > 1,
        "final_scores": dict(scores_by_player), "final_ranks": ranks,
        "agent_scores": {seat_assignment[pid]: s for pid, s in scores_by_player.items()},
        "agent_ranks": {seat_assignment[pid]: r for pid, r in ranks.items()},
        "winner_id": winner_ids[0] if len(winner_ids) == 1 else None, "moves_made": int(moves_made),
        "turn_count": int(turn_count), "passes": int(passes), "invalid_actions": int(invalid_actions),
        "duration_sec": float(duration), "truncated": bool(truncated), "agent_move_stats": per_agent_stats,
        "snapshot_checkpoints_hit": [], "error": error,
    }


# ---- object-API game loop (arena_runner.py:578-777) -----------------------------------------------------
def _choose(agent, board, player, legal_moves, thinking_time_ms) -> Tuple[Optional[Any], Dict[str, Any]]:
    start = time.perf_counter()
    if hasattr(agent, "choose_move"):
        move, stats = agent.choose_move(board, player, legal_moves, int(thinking_time_ms or 1))
        return move, dict(stats or {})
    if hasattr(agent, "think") and thinking_time_ms:
        result = agent.think(board, player, legal_moves, int(thinking_time_ms))
        return result.get("move"), dict(result.get("stats") or {})
    move = agent.select_action(board, player, legal_moves)
    stats: Dict[str, Any] = {"timeSpentMs": (time.perf_counter() - start) * 1000.0}
    info = getattr(agent, "stats", None)
    if isinstance(info, dict) and info.get("iterations_run") is not None:
        stats["iterations_run"] = info["iterations_run"]
    return move, stats


def run_single_game(*, run_id: str, game_index: int, game_seed: int, seat_assignment: Mapping[str, str],
                    agents: Mapping[str, Any], max_turns: int = 2500, thinking_time_ms: Optional[int] = None,
                    seat_policy: str = "round_robin") -> Dict[str, Any]:
    random.seed(game_seed)
    np.random.seed(game_seed)
    start = time.perf_counter()
    stats = {name: {"moves": 0.0, "total_time_ms": 0.0, "total_simulations": 0.0, "moves_with_simulations": 0.0,
                    "move_times_ms": []} for name in set(seat_assignment.values())}
    game = BlokusGame(enable_telemetry=False)
    passes = invalid = turns = 0
    error = None
    try:
        while not game.is_game_over() and turns < max_turns:
            player = game.get_current_player()
            name = seat_assignment[str(player.value)]
            legal = game.get_legal_moves(player)
            turns += 1
            if not legal:
                passes += 1
                game.board._update_current_player()
                game._check_game_over()
                continue
            t0 = time.perf_counter()
            move, raw = _choose(agents[name], game.board, player, legal, thinking_time_ms)
            elapsed = float(raw.get("timeSpentMs", (time.perf_counter() - t0) * 1000.0))
            sims = raw.get("nodesEvaluated", raw.get("iterations_run"))
            st = stats[name]
            st["moves"] += 1
            st["total_time_ms"] += elapsed
            st["move_times_ms"].append(elapsed)
            if sims is not None:
                st["total_simulations"] += sims
                st["moves_with_simulations"] += 1
            if move is None or not game.make_move(move, player):
                invalid += 0 if move is None else 1
                passes += 1
                game.board._update_current_player()
                game._check_game_over()
    except Exception:
        error = traceback.format_exc()
    truncated = turns >= max_turns and not game.is_game_over()
    if truncated:
        game.board.game_over = True
    for st in stats.values():
        st["avg_time_ms"] = st["total_time_ms"] / st["moves"] if st["moves"] else 0.0
        if st["moves_with_simulations"] > 0:
            st["avg_simulations_per_move"] = st["total_simulations"] / st["moves_with_simulations"]
            secs = st["total_time_ms"] / 1000.0
            st["simulations_per_second"] = st["total_simulations"] / secs if secs > 0 else None
        else:
            st["avg_simulations_per_move"] = st["simulations_per_second"] = st["total_simulations"] = None
    result = game.get_game_result()
    scores = {str(pid): int(s) for pid, s in result.scores.items()}
    return _record(run_id, game_index, game_seed, seat_assignment, seat_policy, scores, game.board.move_count, turns,
                   passes, invalid, time.perf_counter() - start, truncated, stats, error)


# ---- device-batched match-ups between the built-in policies ----------------------------------------------
def run_device_arena(agent_policies: Mapping[str, str], num_games: int, run_seed: int = 42,
                     seat_policy: str = "round_robin", run_id: str = "device_arena",
                     snapshots: Optional[Sequence[int]] = None, max_turns: int = 2500) -> Dict[str, Any]:
    """agent_policies: four agent names -> "random" | "heuristic".  All num_games games run in one launch.

    Returns {"records": [...], "summary": {agent: {"games", "wins", "win_rate", "mean_score", "mean_rank"}}}.
    Seat assignment and the recorded game_seed follow the reference; the moves of game i come from the Philox
    stream (key = run_seed, game index = i).

    snapshots: checkpoint plies (e.g. winprob.DEFAULT_SNAPSHOT_PLYS).  When given, the recorded games are replayed on
    the device and the win-probability feature rows are captured where arena_runner.py:625-650,698 captures them;
    "snapshot_rows" holds the rows of snapshots.csv (arena_runner.py:522-575 + the end-of-game labels, :722-731), with
    phase_progress_* computed against `max_turns` (RunConfig.max_turns).
    """
    import torch

    from . import device as dev

    names = list(agent_policies)
    if len(names) != 4:
        raise ValueError("a Blokus game seats exactly four agents")
    seats, packed = [], np.zeros(num_games, dtype=np.int32)
    for i in range(num_games):
        gs = game_seed_from_run_seed(run_seed, i)
        sa = seat_assignment_for_game(names, i, gs, seat_policy)
        seats.append((gs, sa))
        word = 0
        for p in range(4):
            word |= POLICY_OF[agent_policies[sa[str(p + 1)]]] << (8 * p)
        packed[i] = word
    t0 = time.perf_counter()
    r = dev.playout(num_games, seed=run_seed, seat_policies=torch.from_numpy(packed).cuda(), want_trace=True,
                    want_final=True)
    scores = r["game_scores"].cpu().numpy()
    trace = r["trace"].cpu().numpy()
    finals = r["final_states"].cpu().numpy()
    duration = time.perf_counter() - t0
    records = []
    agg = {n: {"games": 0, "wins": 0.0, "score_sum": 0, "rank_sum": 0} for n in names}
    for i, (gs, sa) in enumerate(seats):
        by_player = {str(p + 1): int(scores[i, p]) for p in range(4)}
        placed = trace[i] != 0xFFFF
        last = int(np.nonzero(placed)[0].max()) + 1 if placed.any() else 0
        rec = _record(run_id, i, gs, sa, seat_policy, by_player, int(placed.sum()), last, int((~placed[:last]).sum()), 0,
                      duration / max(num_games, 1), False, {}, None)
        rec["moves_made"] = int(finals[i, 242]) | (int(finals[i, 243]) << 8)
        records.append(rec)
        for pid, name in sa.items():
            a = agg[name]
            a["games"] += 1
            a["score_sum"] += by_player[pid]
            a["rank_sum"] += rec["final_ranks"][pid]
            if int(pid) in rec["winner_ids"]:
                a["wins"] += 1.0 / len(rec["winner_ids"])
    summary = {n: {"games": a["games"], "wins": a["wins"], "win_rate": a["wins"] / max(a["games"], 1),
                   "mean_score": a["score_sum"] / max(a["games"], 1), "mean_rank": a["rank_sum"] / max(a["games"], 1)}
               for n, a in agg.items()}
    out = {"records": records, "summary": summary, "seat_policies": packed}
    if snapshots:
        from . import winprob

        replay = dev.replay_snapshots(r["trace"], snapshots, max_turns=max_turns)
        hit = (replay["turn_index"] >= 0).cpu().numpy()
        for i, rec in enumerate(records):
            rec["snapshot_checkpoints_hit"] = [int(c) for c, h in zip(replay["checkpoints"], hit[i]) if h]
        out["snapshot_rows"] = winprob.snapshot_rows_from_replay(replay, records)
        out["snapshot_features"] = replay
    return out


def write_games_jsonl(path: str, records: Sequence[Mapping[str, Any]]) -> None:
    with open(path, "w", encoding="utf-8") as f:
        for rec in records:
            f.write(json.dumps(rec, sort_keys=True) + "\n")

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
                    agents: Mapping[str, Any], max_turns: int = 2500, thinking_time_ms: Any = None,
                    seat_policy: str = "round_robin", record_trace: bool = False) -> Dict[str, Any]:
    """thinking_time_ms: one budget for every agent or a mapping agent name -> budget (AgentConfig.thinking_time_ms).
    record_trace adds "move_trace" (move id per turn, 0xFFFF = pass) to the record, the input of
    device.replay_snapshots."""
    random.seed(game_seed)
    np.random.seed(game_seed)
    start = time.perf_counter()
    stats = {name: {"moves": 0.0, "total_time_ms": 0.0, "total_simulations": 0.0, "moves_with_simulations": 0.0,
                    "move_times_ms": []} for name in set(seat_assignment.values())}
    game = BlokusGame(enable_telemetry=False)
    passes = invalid = turns = 0
    error = None
    trace = []
    try:
        while not game.is_game_over() and turns < max_turns:
            player = game.get_current_player()
            name = seat_assignment[str(player.value)]
            legal = game.get_legal_moves(player)
            turns += 1
            trace.append(0xFFFF)
            if not legal:
                passes += 1
                game.board._update_current_player()
                game._check_game_over()
                continue
            t0 = time.perf_counter()
            budget = thinking_time_ms.get(name) if isinstance(thinking_time_ms, Mapping) else thinking_time_ms
            move, raw = _choose(agents[name], game.board, player, legal, budget)
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
            else:
                trace[-1] = move.to_id()
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
    rec = _record(run_id, game_index, game_seed, seat_assignment, seat_policy, scores, game.board.move_count, turns,
                  passes, invalid, time.perf_counter() - start, truncated, stats, error)
    if record_trace:
        rec["move_trace"] = trace
    return rec


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


# ---- experiment configuration and driver (arena_runner.py:41-236, 308-492, 914-1000) -------------------------------
DEFAULT_OUTPUT_ROOT = "arena_runs"
DEFAULT_MAX_TURNS = 2500
SUPPORTED_SEAT_POLICIES = {"randomized", "round_robin"}
DEFAULT_SNAPSHOT_PLYS = [8, 16, 24, 32, 40, 48, 56, 64]


class AgentConfig:
    """One arena agent: the reference's JSON shape {"name", "type", "thinking_time_ms", "params"} (arena_runner.py:52-87);
    keys outside these four are folded into params."""

    def __init__(self, name: str, type: str, thinking_time_ms: Optional[int] = None, params: Optional[Dict[str, Any]] = None):
        self.name, self.type = str(name), str(type)
        self.thinking_time_ms = int(thinking_time_ms) if thinking_time_ms is not None else None
        self.params = dict(params or {})

    @classmethod
    def from_dict(cls, item: Mapping[str, Any]) -> "AgentConfig":
        if "name" not in item:
            raise ValueError("Agent entries must include 'name'")
        if "type" not in item:
            raise ValueError(f"Agent '{item['name']}' is missing required field 'type'")
        params = dict(item.get("params") or {})
        params.update({k: v for k, v in item.items() if k not in {"name", "type", "thinking_time_ms", "params"}})
        return cls(item["name"], item["type"], item.get("thinking_time_ms"), params)

    def to_dict(self) -> Dict[str, Any]:
        return {"name": self.name, "type": self.type, "thinking_time_ms": self.thinking_time_ms, "params": dict(self.params)}


class SnapshotConfig:
    """arena_runner.py:90-121."""

    def __init__(self, enabled: bool = False, strategy: str = "fixed_ply", checkpoints: Optional[Sequence[int]] = None):
        self.enabled, self.strategy = bool(enabled), str(strategy)
        self.checkpoints = sorted({int(v) for v in (DEFAULT_SNAPSHOT_PLYS if checkpoints is None else checkpoints) if int(v) >= 0})

    @classmethod
    def from_dict(cls, item: Mapping[str, Any]) -> "SnapshotConfig":
        cps = item.get("checkpoints", list(DEFAULT_SNAPSHOT_PLYS))
        if not isinstance(cps, list):
            raise ValueError("snapshots.checkpoints must be a list of integers")
        return cls(item.get("enabled", False), item.get("strategy", "fixed_ply"), cps)

    def validate(self) -> None:
        if self.strategy != "fixed_ply":
            raise ValueError(f"Unsupported snapshot strategy '{self.strategy}'. Supported values: ['fixed_ply'].")

    def to_dict(self) -> Dict[str, Any]:
        return {"enabled": self.enabled, "strategy": self.strategy, "checkpoints": list(self.checkpoints)}


class RunConfig:
    """Top-level experiment configuration with the reference's keys, defaults and validation (arena_runner.py:124-236),
    including the legacy {name: {type, ...}} shape of scripts/arena_config.json."""

    def __init__(self, agents: Sequence[AgentConfig], num_games: int, seed: int, seat_policy: str = "randomized",
                 output_root: str = DEFAULT_OUTPUT_ROOT, max_turns: int = DEFAULT_MAX_TURNS, notes: str = "",
                 snapshots: Optional[SnapshotConfig] = None):
        self.agents, self.num_games, self.seed = list(agents), int(num_games), int(seed)
        self.seat_policy, self.output_root, self.max_turns, self.notes = str(seat_policy), str(output_root), int(max_turns), str(notes)
        self.snapshots = snapshots or SnapshotConfig()

    @classmethod
    def from_dict(cls, config: Mapping[str, Any]) -> "RunConfig":
        raw = config.get("agents")
        if raw is None:                                          # legacy shape
            raw = []
            for name, item in config.items():
                if not isinstance(item, Mapping) or "type" not in item:
                    continue
                ttm = item.get("thinking_time_ms")
                if ttm is None and item.get("time_limit") is not None:
                    ttm = int(float(item["time_limit"]) * 1000)
                raw.append({"name": str(name), "type": str(item["type"]), "thinking_time_ms": ttm,
                            "params": {k: v for k, v in item.items() if k not in {"type", "thinking_time_ms"}}})
            if not raw:
                raise ValueError("Invalid run config: missing 'agents' list and no legacy agent entries were found.")
        if not isinstance(raw, list):
            raise ValueError("RunConfig 'agents' must be a list.")
        snaps = config.get("snapshots") or {}
        if not isinstance(snaps, Mapping):
            raise ValueError("RunConfig 'snapshots' must be an object when provided.")
        cfg = cls([AgentConfig.from_dict(a) for a in raw], config.get("num_games", 100), config.get("seed", 0),
                  config.get("seat_policy", "randomized"), config.get("output_root", DEFAULT_OUTPUT_ROOT),
                  config.get("max_turns", DEFAULT_MAX_TURNS), config.get("notes", ""), SnapshotConfig.from_dict(snaps))
        cfg.validate()
        return cfg

    def validate(self) -> None:
        if self.num_games <= 0:
            raise ValueError("num_games must be > 0.")
        if len(self.agents) != 4:
            raise ValueError(f"Arena expects exactly 4 agents; received {len(self.agents)}.")
        if len({a.name for a in self.agents}) != len(self.agents):
            raise ValueError("Agent names must be unique.")
        if self.seat_policy not in SUPPORTED_SEAT_POLICIES:
            raise ValueError(f"Unsupported seat_policy '{self.seat_policy}'. Expected one of {sorted(SUPPORTED_SEAT_POLICIES)}.")
        if self.max_turns <= 0:
            raise ValueError("max_turns must be > 0.")
        self.snapshots.validate()

    def to_dict(self) -> Dict[str, Any]:
        return {"agents": [a.to_dict() for a in self.agents], "num_games": self.num_games, "seed": self.seed,
                "seat_policy": self.seat_policy, "output_root": self.output_root, "max_turns": self.max_turns,
                "notes": self.notes, "snapshots": self.snapshots.to_dict()}

    @property
    def agent_names(self):
        return [a.name for a in self.agents]


def load_run_config(config_path: str) -> RunConfig:
    with open(config_path, "r", encoding="utf-8") as f:
        payload = json.load(f)
    if not isinstance(payload, Mapping):
        raise ValueError("Run config must be a JSON object.")
    return RunConfig.from_dict(payload)


class _BudgetedAgent:
    """choose_move wrapper for the think-style agents with the reference's deterministic budget rule
    (arena_runner.py:334-412): iteration cap = round(iterations_per_ms * thinking_time_ms), wall clock not limiting."""

    def __init__(self, agent, deterministic: bool, iterations_per_ms: float):
        self.agent, self.deterministic, self.iterations_per_ms = agent, deterministic, iterations_per_ms

    def choose_move(self, board, player, legal_moves, thinking_time_ms):
        inner = getattr(self.agent, "_agent", self.agent)                    # GameplayFastMCTSAgent wraps a FastMCTSAgent
        budget = int(thinking_time_ms or max(int(getattr(inner, "time_limit", 0.001) * 1000), 1))
        if not self.deterministic:
            res = inner.think(board, player, legal_moves, budget)
            return res.get("move"), dict(res.get("stats") or {})
        cap, keep = max(1, int(round(self.iterations_per_ms * budget))), inner.iterations
        inner.iterations = cap
        try:
            res = inner.think(board, player, legal_moves, max(10_000_000, budget))
        finally:
            inner.iterations = keep
        stats = dict(res.get("stats") or {})
        stats["timeBudgetMs"], stats["iterationCap"] = budget, cap
        return res.get("move"), stats


def build_agent(config: AgentConfig, seed: int):
    """The reference's agent factory (arena_runner.py:415-492) over this package's agents; same type names, parameter
    names and defaults.  `leaf_batch` (params) is the one extra knob: warps per tree for the MCTS types."""
    from .agents import FastMCTSAgent, GameplayFastMCTSAgent, HeuristicAgent, RandomAgent
    from .mcts import MCTSAgent

    kind, p = config.type.lower(), dict(config.params)
    if kind == "random":
        return RandomAgent(seed=seed)
    if kind == "heuristic":
        agent = HeuristicAgent(seed=seed)
        if isinstance(p.get("weights"), Mapping):
            agent.set_weights(dict(p["weights"]))
        return agent
    if kind == "mcts":
        det, ipm = bool(p.get("deterministic_time_budget", True)), float(p.get("iterations_per_ms", 10.0))
        iterations, time_limit = int(p.get("iterations", 1000)), p.get("time_limit")
        if det and config.thinking_time_ms is not None:
            iterations, time_limit = max(1, int(round(ipm * config.thinking_time_ms))), None
        elif time_limit is None and config.thinking_time_ms is not None:
            time_limit = float(config.thinking_time_ms) / 1000.0
        leaf_batch = int(p.get("leaf_batch", 1))
        return MCTSAgent(
            iterations=iterations, time_limit=float(time_limit) if time_limit is not None else None,
            exploration_constant=float(p.get("exploration_constant", 1.414)),
            use_transposition_table=bool(p.get("use_transposition_table", True)), seed=seed,
            learned_model_path=p.get("learned_model_path"), leaf_evaluation_enabled=bool(p.get("leaf_evaluation_enabled", False)),
            progressive_bias_enabled=bool(p.get("progressive_bias_enabled", False)),
            progressive_bias_weight=float(p.get("progressive_bias_weight", 0.25)),
            potential_shaping_enabled=bool(p.get("potential_shaping_enabled", False)),
            potential_shaping_gamma=float(p.get("potential_shaping_gamma", 1.0)),
            potential_shaping_weight=float(p.get("potential_shaping_weight", 1.0)),
            potential_mode=str(p.get("potential_mode", "prob")), max_rollout_moves=int(p.get("max_rollout_moves", 50)),
            leaf_batch=leaf_batch)
    if kind == "fast_mcts":
        default_time = float(config.thinking_time_ms) / 1000.0 if config.thinking_time_ms is not None else float(p.get("time_limit", 0.1))
        agent = FastMCTSAgent(iterations=int(p.get("iterations", 5000)), time_limit=default_time,
                              exploration_constant=float(p.get("exploration_constant", 1.414)), seed=seed,
                              leaf_batch=int(p.get("leaf_batch", 1)))
        return _BudgetedAgent(agent, bool(p.get("deterministic_time_budget", True)), float(p.get("iterations_per_ms", 20.0)))
    if kind in {"gameplay_fast_mcts", "gameplay_mcts"}:
        agent = GameplayFastMCTSAgent(iterations=int(p.get("iterations", 5000)),
                                      exploration_constant=float(p.get("exploration_constant", 1.414)), seed=seed)
        return _BudgetedAgent(agent, bool(p.get("deterministic_time_budget", True)), float(p.get("iterations_per_ms", 20.0)))
    raise ValueError(f"Unsupported agent type: {config.type}")


def summarize(records: Sequence[Mapping[str, Any]], agent_names: Sequence[str]) -> Dict[str, Any]:
    """Per-agent games / wins (ties split) / win rate / mean score / mean rank.  The reference's full statistics module
    (pairwise tables, Elo, confidence intervals: analytics/tournament/arena_stats.py) is outside the hot path."""
    agg = {n: {"games": 0, "wins": 0.0, "score_sum": 0, "rank_sum": 0} for n in agent_names}
    done = 0
    for rec in records:
        if rec.get("error"):
            continue
        done += 1
        for name, score in rec["agent_scores"].items():
            a = agg[name]
            a["games"] += 1
            a["score_sum"] += score
            a["rank_sum"] += rec["agent_ranks"][name]
            if name in rec["winner_agents"]:
                a["wins"] += 1.0 / len(rec["winner_agents"])
    return {"completed_games": done, "error_games": len(records) - done,
            "agents": {n: {"games": a["games"], "wins": a["wins"], "win_rate": a["wins"] / max(a["games"], 1),
                           "mean_score": a["score_sum"] / max(a["games"], 1), "mean_rank": a["rank_sum"] / max(a["games"], 1)}
                       for n, a in agg.items()}}


def run_experiment(run_config: RunConfig, *, verbose: bool = False, device_batch: bool = True) -> Dict[str, Any]:
    """arena_runner.py:914-1000: run all games of a RunConfig and write run_config.json, games.jsonl, snapshots.csv
    (+ .parquet) and summary.json under <output_root>/<run_id>/.

    When every agent is of a built-in policy type ("random" / "heuristic", default weights) and `device_batch` is set,
    the whole experiment is ONE device batch (run_device_arena): same seeds, seat assignments, record keys and
    snapshot rows, moves drawn with Philox.  Mixes that include search agents ("mcts", "fast_mcts",
    "gameplay_fast_mcts" -- the reference's own arena configs) run as device batches too, one search launch per agent and
    turn over all games (run_device_arena_agents).  Anything else (custom heuristic weights, learned models,
    device_batch=False) runs the reference's per-game loop over agent objects (a fresh agent per game and name, seeded
    with the reference's agent_seed)."""
    import os
    from datetime import datetime

    from . import winprob

    stamp = datetime.now()
    digest = hashlib.sha256(json.dumps(run_config.to_dict(), sort_keys=True).encode("utf-8")).hexdigest()[:8]
    base_id = f"{stamp.strftime('%Y%m%d_%H%M%S')}_{digest}"
    os.makedirs(run_config.output_root, exist_ok=True)
    run_id, attempt = base_id, 1
    while os.path.exists(os.path.join(run_config.output_root, run_id)):
        run_id, attempt = f"{base_id}_{attempt:02d}", attempt + 1
    run_dir = os.path.join(run_config.output_root, run_id)
    os.makedirs(run_dir)
    payload = dict(run_config.to_dict(), run_id=run_id, created_at=stamp.isoformat(timespec="seconds"))
    with open(os.path.join(run_dir, "run_config.json"), "w", encoding="utf-8") as f:
        json.dump(payload, f, indent=2, sort_keys=True)
        f.write("\n")
    builtin = all(a.type.lower() in POLICY_OF and not a.params.get("weights") for a in run_config.agents)
    batchable = all((a.type.lower() in POLICY_OF and not a.params.get("weights")) or
                    (a.type.lower() in _SEARCH_TYPES and not a.params.get("learned_model_path")) for a in run_config.agents)
    snaps = run_config.snapshots.checkpoints if run_config.snapshots.enabled else None
    if batchable and not builtin and device_batch:
        out = run_device_arena_agents(run_config, run_id=run_id)
        records, rows = out["records"], out.get("snapshot_rows", [])
        for rec in records:
            rec["agent_move_stats"] = {n: out["agent_move_stats"][n] for n in set(rec["seat_assignment"].values())}
    elif builtin and device_batch:
        out = run_device_arena({a.name: a.type.lower() for a in run_config.agents}, run_config.num_games,
                               run_seed=run_config.seed, seat_policy=run_config.seat_policy, run_id=run_id,
                               snapshots=snaps, max_turns=run_config.max_turns)
        records, rows = out["records"], out.get("snapshot_rows", [])
    else:
        records, rows = [], []
        cfg_of = {a.name: a for a in run_config.agents}
        for gi in range(run_config.num_games):
            gs = game_seed_from_run_seed(run_config.seed, gi)
            seats = seat_assignment_for_game(run_config.agent_names, gi, gs, run_config.seat_policy)
            agents = {n: build_agent(cfg_of[n], agent_seed(run_config.seed, gi, n)) for n in set(seats.values())}
            think_ms = {n: cfg_of[n].thinking_time_ms for n in agents}
            rec = run_single_game(run_id=run_id, game_index=gi, game_seed=gs, seat_assignment=seats, agents=agents,
                                  max_turns=run_config.max_turns, thinking_time_ms=think_ms,
                                  seat_policy=run_config.seat_policy, record_trace=bool(snaps))
            records.append(rec)
            if verbose:
                print(f"[{gi + 1}/{run_config.num_games}] seed={gs} winners={','.join(rec['winner_agents'])} scores={rec['agent_scores']}")
        if snaps:                                     # replay the recorded moves on the device, capture the feature rows
            import torch

            from . import device as dev

            traces = [rec.pop("move_trace") for rec in records]
            width = max(96, max(len(t) for t in traces))
            mat = np.full((len(traces), width), 0xFFFF, dtype=np.uint16)
            for i, t in enumerate(traces):
                mat[i, :len(t)] = t
            replay = dev.replay_snapshots(torch.from_numpy(mat).cuda(), snaps, max_turns=run_config.max_turns)
            hit = (replay["turn_index"] >= 0).cpu().numpy()
            for i, rec in enumerate(records):
                rec["snapshot_checkpoints_hit"] = [int(c) for c, h in zip(replay["checkpoints"], hit[i]) if h]
            rows = winprob.snapshot_rows_from_replay(replay, records)
    write_games_jsonl(os.path.join(run_dir, "games.jsonl"), records)
    summary = dict(summarize(records, run_config.agent_names), run_id=run_id, run_seed=run_config.seed,
                   seat_policy=run_config.seat_policy, run_config=payload,
                   thinking_time_ms_by_agent={a.name: a.thinking_time_ms for a in run_config.agents})
    meta = winprob.write_snapshots_dataset(run_dir, rows)
    if meta["enabled"]:
        summary["snapshots"] = dict(meta, expected_rows_max=run_config.num_games * len(run_config.snapshots.checkpoints) * 4)
    with open(os.path.join(run_dir, "summary.json"), "w", encoding="utf-8") as f:
        json.dump(summary, f, indent=2, sort_keys=True)
        f.write("\n")
    return {"run_id": run_id, "run_dir": run_dir, "summary": summary, "snapshot_rows": len(rows)}


# ---- whole experiments with search agents as device batches ------------------------------------------------------------
_SEARCH_TYPES = {"mcts", "fast_mcts", "gameplay_fast_mcts", "gameplay_mcts"}


def _iteration_cap(cfg: AgentConfig) -> int:
    """Simulations per move of a search agent under the reference's deterministic budget rule (arena_runner.py:334-356,
    446-450): round(iterations_per_ms * thinking_time_ms), else its `iterations` parameter."""
    p = cfg.params
    kind = cfg.type.lower()
    ipm = float(p.get("iterations_per_ms", 10.0 if kind == "mcts" else 20.0))
    default_iters = int(p.get("iterations", 1000 if kind == "mcts" else 5000))
    if cfg.thinking_time_ms is not None and bool(p.get("deterministic_time_budget", True)):
        return max(1, int(round(ipm * cfg.thinking_time_ms)))
    if cfg.thinking_time_ms is not None:                    # wall-clock budgets have no meaning inside one batch
        return max(1, min(default_iters, int(round(ipm * cfg.thinking_time_ms))))
    return max(1, default_iters)


def run_device_arena_agents(run_config: RunConfig, run_id: str = "device_arena", auto_leaf_batch: bool = True) -> Dict[str, Any]:
    """All games of a RunConfig at once on the device, for any mix of "random", "heuristic", "mcts", "fast_mcts" and
    "gameplay_fast_mcts" agents (the reference's scripts/arena_config.json seats four gameplay_fast_mcts agents with
    25 / 50 / 100 / 200 ms budgets).  Per turn: one single-turn playout launch draws the moves of the built-in policies
    (same Philox streams as run_device_arena), one search launch per search agent covers all games where that agent is
    to move (flat UCB bandit with 50-ply rollouts for the fast types, UCT with HeuristicAgent rollouts for "mcts",
    simulations per move from the deterministic budget rule), then one bk_step applies every move.  Seeds, seat
    assignments, record keys and snapshot rows are those of run_experiment.

    auto_leaf_batch: a flat-bandit agent (fast types) without an explicit "leaf_batch" parameter gets leaf-batched rollouts (up to 16
    warps per tree) when the games alone would leave most of the GPU idle -- about a quarter of the games are searched
    per launch, and one warp per tree fills the device only from ~14 000 games on.  Its budget stays the same number
    of simulations per move."""
    import torch

    from . import device as dev
    from .batched import PASS, BatchedBlokus
    from .mcts import MAX_ROOT_MOVES, DeviceUCT

    names, n = run_config.agent_names, run_config.num_games
    cfgs = list(run_config.agents)
    for a in cfgs:
        if a.type.lower() not in POLICY_OF and a.type.lower() not in _SEARCH_TYPES:
            raise ValueError(f"Unsupported agent type for the device arena: {a.type}")
    seats, agent_at = [], np.zeros((n, 4), dtype=np.int64)
    packed = np.zeros(n, dtype=np.int32)
    for i in range(n):
        gs = game_seed_from_run_seed(run_config.seed, i)
        sa = seat_assignment_for_game(names, i, gs, run_config.seat_policy)
        seats.append((gs, sa))
        for p in range(4):
            k = names.index(sa[str(p + 1)])
            agent_at[i, p] = k
            packed[i] |= POLICY_OF.get(cfgs[k].type.lower(), 0) << (8 * p)       # search seats: placeholder policy
    t0 = time.perf_counter()
    games = BatchedBlokus(n)
    device = games.states.device
    agent_at_d = torch.from_numpy(agent_at).to(device)
    packed_d = torch.from_numpy(packed).to(device)
    engines: Dict[int, Any] = {}
    caps = {k: _iteration_cap(c) for k, c in enumerate(cfgs) if c.type.lower() in _SEARCH_TYPES}
    for k, c in enumerate(cfgs):
        if k in caps:
            flat = c.type.lower() != "mcts"
            lb = int(c.params.get("leaf_batch", 0))
            if lb <= 0:                          # only the flat bandits: batching changes how a UCT tree grows
                lb = min(16, max(1, 3552 // max(1, -(-n // 4)))) if (auto_leaf_batch and flat) else 1
            iters = -(-caps[k] // lb)
            engines[k] = (DeviceUCT(iters, n_trees=n, max_rollout_moves=int(c.params.get("max_rollout_moves", 50)),
                                    max_depth=1 if flat else 0, ucb_variant=1 if flat else 0, reward_mode=1 if flat else 0,
                                    exploration=float(c.params.get("exploration_constant", 1.414)),
                                    seed=stable_hash_int(run_config.seed, c.name, "device_arena"),
                                    rollout_policy=dev.POLICY_UNIFORM if flat else dev.POLICY_HEURISTIC, leaf_batch=lb), iters, lb)
    max_t = 96
    trace = torch.full((n, max_t), PASS, dtype=torch.int32, device=device)
    turns = torch.zeros(n, dtype=torch.int32, device=device)
    moves_by = np.zeros(len(cfgs), dtype=np.int64)
    rank_weight = torch.arange(MAX_ROOT_MOVES, device=device, dtype=torch.int64)
    scratch: Dict[str, Any] = {}
    for t in range(max_t):
        running = games.alive_mask() != 0
        if not bool(running.any()):
            break
        mover = games.states[:, 240].to(torch.int64)
        who = agent_at_d.gather(1, mover[:, None])[:, 0]
        counts = games.legal_move_counts()
        can = running & (counts > 0)
        r = dev.playout(n, seed=run_config.seed, start_states=games.states, seat_policies=packed_d, first_turn=t,
                        max_turns=t + 1, want_trace=True, want_board_scores=False, out=scratch)
        move = r["trace"][:, t].to(torch.int32)                                  # built-in seats (and placeholders)
        for k, (eng, iters, lb) in engines.items():
            idx = torch.nonzero(can & (who == k))[:, 0]
            m = int(idx.numel())
            if m == 0:
                continue
            eng.p.tree_offset = t * (1 << 24)
            res = eng.search(games.states[idx].contiguous(), iters, active_trees=m)
            visits = res["visits"][:m]
            key = visits.to(torch.int64) * MAX_ROOT_MOVES + rank_weight[None, : visits.shape[1]]
            best = key.argmax(dim=1)                                             # most visited, ties to the highest rank
            ids = res["moves"][:m].view(torch.int16).gather(1, best[:, None]).view(torch.uint16)[:, 0]
            move[idx] = ids.to(torch.int32)
            moves_by[k] += m
        for k, c in enumerate(cfgs):
            if k not in engines:
                moves_by[k] += int((can & (who == k)).sum())
        move = torch.where(can, move, torch.full_like(move, PASS))
        games.step(move.to(torch.int16).view(torch.uint16).contiguous())
        trace[:, t] = torch.where(running, move, trace[:, t])
        turns += running.to(torch.int32)
    duration = time.perf_counter() - t0
    gsc, _ = games.scores()
    scores = gsc.cpu().numpy()
    tr = trace.cpu().numpy()
    nturns = turns.cpu().numpy()
    finals = games.states.cpu().numpy()
    per_agent = {c.name: {"moves": float(moves_by[k]), "total_simulations": float(moves_by[k] * engines[k][1] * engines[k][2]) if k in engines else None,
                          "avg_simulations_per_move": float(engines[k][1] * engines[k][2]) if k in engines else None}
                 for k, c in enumerate(cfgs)}
    records = []
    for i, (gs, sa) in enumerate(seats):
        by_player = {str(p + 1): int(scores[i, p]) for p in range(4)}
        placed = tr[i, : nturns[i]] != PASS
        rec = _record(run_id, i, gs, sa, run_config.seat_policy, by_player, int(placed.sum()), int(nturns[i]),
                      int((~placed).sum()), 0, duration / max(n, 1), False, {}, None)
        rec["moves_made"] = int(finals[i, 242]) | (int(finals[i, 243]) << 8)
        records.append(rec)
    out = {"records": records, "summary": summarize(records, names), "agent_move_stats": per_agent,
           "trace": trace.to(torch.int16).view(torch.uint16)}
    if run_config.snapshots.enabled:
        from . import winprob

        replay = dev.replay_snapshots(out["trace"], run_config.snapshots.checkpoints, max_turns=run_config.max_turns)
        hit = (replay["turn_index"] >= 0).cpu().numpy()
        for i, rec in enumerate(records):
            rec["snapshot_checkpoints_hit"] = [int(c) for c, h in zip(replay["checkpoints"], hit[i]) if h]
        out["snapshot_rows"] = winprob.snapshot_rows_from_replay(replay, records)
        out["snapshot_features"] = replay
    return out

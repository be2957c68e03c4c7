"""Times the REFERENCE'S OWN Python engine (oracle/_ref, built by oracle/make_ref.py) on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by bench.py (``cpu_baseline`` and ``--impl reference``) and by tests.
Nothing here touches the product package; everything below calls the reference's public API only:

  * playouts      -- the arena game loop (analytics/tournament/arena_runner.py:653-695) on
                     ``BlokusGame(enable_telemetry=False)`` with a uniform ``random.Random`` choice
                     (agents/random_agent.py:33-50), one process per core;
  * movegen       -- the protocol of benchmarks/benchmark_move_generation.py:136-224 with the benchmark's own timing
                     functions (its five seeded states, frontier+bitboard = the default generator, and the naive one);
  * search        -- ``MCTSAgent`` (uniform and default heuristic rollouts) and ``FastMCTSAgent.think`` on one core.
"""

from __future__ import annotations

import os
import random
import sys
import time
from typing import Dict, List, Tuple

_READY = False


def _ensure_reference() -> None:
    global _READY
    if _READY:
        return
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    from oracle import make_ref

    tree = make_ref.ref_path()
    if tree not in sys.path:
        sys.path.insert(0, tree)
    sys.dont_write_bytecode = True
    import logging

    logging.disable(logging.WARNING)          # the engine logs per move when it is slow
    _READY = True


def play_games(seed: int, n_games: int) -> Tuple[int, int, float, List[int]]:
    """n_games uniform-random games to the end; (games, plies, seconds, scores of the last game)."""
    _ensure_reference()
    from engine.game import BlokusGame

    rng = random.Random(seed)
    t0 = time.perf_counter()
    plies = 0
    last = []
    for _ in range(n_games):
        game = BlokusGame(enable_telemetry=False)
        turns = 0
        while not game.is_game_over() and turns < 2500:
            player = game.get_current_player()
            legal = game.get_legal_moves(player)
            if not legal:                                  # arena_runner.py:662-666
                game.board._update_current_player()
                game._check_game_over()
                turns += 1
                continue
            game.make_move(rng.choice(legal), player)      # agents/random_agent.py:47-50
            plies += 1
            turns += 1
        res = game.get_game_result()
        last = [int(res.scores[p]) for p in sorted(res.scores)]
    return n_games, plies, time.perf_counter() - t0, last


def _init_worker() -> None:
    _ensure_reference()
    import engine.game  # noqa: F401  (import cost outside the timed region)


class PlayoutPool:
    """One process per core, reference imported once per process; ``step(games_per_worker)`` plays a bounded sample."""

    def __init__(self, workers: int):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor

        self.workers = workers
        self.ex = ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn"), initializer=_init_worker)
        self._k = 0
        list(self.ex.map(_noop, range(workers * 2)))       # start every worker now

    def step(self, games_per_worker: int = 1) -> Tuple[int, float]:
        """(games played, wall seconds) -- all workers in parallel."""
        t0 = time.perf_counter()
        seeds = [1_000_003 * (self._k + 1) + w for w in range(self.workers)]
        self._k += 1
        done = list(self.ex.map(play_games, seeds, [games_per_worker] * self.workers))
        return sum(d[0] for d in done), time.perf_counter() - t0

    def close(self) -> None:
        self.ex.shutdown(wait=True, cancel_futures=True)


def _noop(_):
    _ensure_reference()
    return os.getpid()


def movegen_benchmark(num_runs: int = 10) -> Dict:
    """benchmarks/benchmark_move_generation.py's protocol (its states, its timing functions); mean ms per state."""
    _ensure_reference()
    import benchmarks.benchmark_move_generation as bm
    from engine.move_generator import LegalMoveGenerator
    from tests.utils_game_states import generate_random_valid_state

    gen = LegalMoveGenerator()
    rows = []
    for n in (5, 10, 20, 30, 40):
        board, player = generate_random_valid_state(n, seed=n)
        fb, n_moves, _ = bm.benchmark_frontier_bitboard_generator(gen, board, player, num_runs)
        fg, _, _ = bm.benchmark_frontier_grid_generator(gen, board, player, max(2, num_runs // 2))
        nv, _ = bm.benchmark_naive_generator(gen, board, player, 2)
        rows.append({"plies": n, "legal_moves": n_moves, "frontier_bitboard_ms": fb, "frontier_grid_ms": fg, "naive_ms": nv})
    mean_fb = sum(r["frontier_bitboard_ms"] for r in rows) / len(rows)
    return {"states": rows, "frontier_bitboard_gens_per_sec": 1e3 / mean_fb,
            "naive_gens_per_sec": 1e3 / (sum(r["naive_ms"] for r in rows) / len(rows)), "runs_per_state": num_runs}


def search_rates(budget_s: float = 4.0) -> Dict:
    """MCTSAgent simulations/s (uniform and default heuristic rollouts) and FastMCTSAgent on one core, empty board."""
    _ensure_reference()
    from agents.fast_mcts_agent import FastMCTSAgent
    from agents.random_agent import RandomAgent
    from engine.game import BlokusGame
    from mcts.mcts_agent import MCTSAgent

    game = BlokusGame(enable_telemetry=False)
    player = game.get_current_player()
    legal = game.get_legal_moves(player)
    out = {}
    for key, rollout in (("mcts_agent_uniform_rollout_sims_per_sec", RandomAgent(seed=1)), ("mcts_agent_heuristic_rollout_sims_per_sec", None)):
        agent = MCTSAgent(iterations=10 ** 9, time_limit=budget_s, rollout_agent=rollout, seed=1)
        t0 = time.perf_counter()
        agent.select_action(game.board, player, legal)
        out[key] = agent.stats["iterations_run"] / (time.perf_counter() - t0)
    fast = FastMCTSAgent(iterations=1000, time_limit=10.0, seed=1)
    t0 = time.perf_counter()
    res = fast.think(game.board, player, legal, 10_000)
    dt = time.perf_counter() - t0
    out["fast_mcts_1000_iterations_ms"] = dt * 1e3
    out["fast_mcts_pseudo_sims_per_sec"] = float(res["stats"].get("nodesEvaluated", 1000)) / dt
    return out

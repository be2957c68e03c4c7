"""Agents with the reference's interfaces: ``select_action(board, player, legal_moves)``, ``think``,
``choose_move``, ``reset``, ``set_seed``, ``get_action_info``.

  RandomAgent            agents/random_agent.py:14-66
  HeuristicAgent         agents/heuristic_agent.py:14-282   (move scores on the device, softmax draw as in the reference)
  FastMCTSAgent          agents/fast_mcts_agent.py:78-338   (flat UCB bandit over root moves; real device rollouts)
  GameplayFastMCTSAgent  agents/gameplay_fast_mcts.py:14-37
"""

from __future__ import annotations

import math
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import device as dev
from .board import Board, Player
from .mcts import DeviceUCT, WideUCT, _root_state
from .move_generator import Move, get_shared_generator


class RandomAgent:
    """Uniform pick with ``np.random.RandomState(seed).randint(0, len(legal_moves))`` (random_agent.py:49)."""

    def __init__(self, seed: Optional[int] = None):
        self.rng = np.random.RandomState(seed)
        self.move_generator = get_shared_generator()

    def select_action(self, board: Board, player: Player, legal_moves: List[Move]) -> Optional[Move]:
        if not legal_moves:
            return None
        return legal_moves[self.rng.randint(0, len(legal_moves))]

    def get_action_info(self) -> Dict[str, Any]:
        return {"name": "RandomAgent", "type": "random", "description": "Selects moves uniformly from legal actions"}

    def reset(self):
        pass

    def set_seed(self, seed: int):
        self.rng = np.random.RandomState(seed)


class HeuristicAgent:
    """Softmax(T=1) sample over 1.0*size + 2.0*corner_creation - 1.5*edge + 0.5*centre.

    The float64 scores come from ``bk_heuristic_scores`` (one launch for the whole move list) and equal the
    reference's ``_evaluate_move`` bit for bit, for the default weights and for any ``set_weights`` (the four weights
    are kernel arguments; products and sums in the reference's order).
    """

    def __init__(self, seed: Optional[int] = None):
        self.rng = np.random.RandomState(seed)
        self.move_generator = get_shared_generator()
        self.piece_size_weight = 1.0
        self.corner_creation_weight = 2.0
        self.edge_avoidance_weight = -1.5
        self.center_preference_weight = 0.5

    def _default_weights(self) -> bool:
        return (self.piece_size_weight, self.corner_creation_weight, self.edge_avoidance_weight,
                self.center_preference_weight) == (1.0, 2.0, -1.5, 0.5)

    def evaluate_moves(self, board: Board, player: Player, legal_moves: List[Move]) -> np.ndarray:
        weights = None if self._default_weights() else (self.piece_size_weight, self.corner_creation_weight,
                                                        self.edge_avoidance_weight, self.center_preference_weight)
        ids = np.array([m.to_id() for m in legal_moves], dtype=np.uint16)
        st = dev.to_device(np.atleast_1d(board.to_packed()))
        players = torch.tensor([player.value - 1], dtype=torch.uint8, device="cuda")
        offsets = torch.tensor([0, len(ids)], dtype=torch.int32, device="cuda")
        scores = dev.heuristic_scores(st, offsets, torch.from_numpy(ids).cuda(), players, weights=weights)
        return scores.cpu().numpy()

    def _evaluate_move(self, board: Board, player: Player, move: Move) -> float:
        return float(self.evaluate_moves(board, player, [move])[0])

    def select_action(self, board: Board, player: Player, legal_moves: List[Move]) -> Optional[Move]:
        if not legal_moves:
            return None
        probabilities = self._softmax(self.evaluate_moves(board, player, legal_moves), temperature=1.0)
        return legal_moves[self.rng.choice(len(legal_moves), p=probabilities)]

    @staticmethod
    def _softmax(x: np.ndarray, temperature: float = 1.0) -> np.ndarray:
        z = x / temperature
        e = np.exp(z - np.max(z))
        return e / np.sum(e)

    def get_action_info(self) -> Dict[str, Any]:
        return {"name": "HeuristicAgent", "type": "heuristic",
                "description": "Strategic agent with piece size, corner, and edge preferences",
                "weights": {"piece_size": self.piece_size_weight, "corner_creation": self.corner_creation_weight,
                            "edge_avoidance": self.edge_avoidance_weight,
                            "center_preference": self.center_preference_weight}}

    def reset(self):
        pass

    def set_seed(self, seed: int):
        self.rng = np.random.RandomState(seed)

    def set_weights(self, weights: Dict[str, float]):
        self.piece_size_weight = weights.get("piece_size", self.piece_size_weight)
        self.corner_creation_weight = weights.get("corner_creation", self.corner_creation_weight)
        self.edge_avoidance_weight = weights.get("edge_avoidance", self.edge_avoidance_weight)
        self.center_preference_weight = weights.get("center_preference", self.center_preference_weight)


def compute_policy_entropy(visits: List[int]) -> float:
    total = sum(visits)
    if total <= 0:
        return 0.0
    ent = 0.0
    for v in visits:
        if v > 0:
            p = v / total
            ent -= p * math.log(p)
    return ent


class FastMCTSAgent:
    """Flat UCB1 bandit over the root moves with ``sqrt(2 ln N / n)`` exploration, as the reference's
    FastMCTSAgent, under the same ``think`` budget rule (time budget AND iteration cap, fast_mcts_agent.py:153-154)
    and returning the same stats / diagnostics schema (:202-226).

    Deliberate change: the reference's pseudo-rollout scores one fixed root move plus
    noise regardless of the selected child; here every simulation is a real device rollout from the selected
    child (uniform random, at most ``rollout_plies`` plies), rewarded with the root player's board-score gain.
    """

    CHUNK = 32   # simulations per launch between wall-clock checks

    def __init__(self, iterations: int = 30, time_limit: float = 0.5, exploration_constant: float = 1.414,
                 seed: Optional[int] = None, enable_diagnostics: bool = False, diagnostics_sample_interval: int = 100,
                 rollout_plies: int = 50, leaf_batch: int = 1, search_width: Optional[int] = None,
                 num_trees: Optional[int] = None):
        self.iterations = iterations
        self.time_limit = time_limit
        self.exploration_constant = exploration_constant
        self.seed = 0 if seed is None else int(seed)
        self.move_generator = get_shared_generator()
        self.enable_diagnostics = enable_diagnostics
        self.diagnostics_sample_interval = max(1, int(diagnostics_sample_interval))
        self.rollout_plies = int(rollout_plies)
        # GPU-batched leaf rollouts: `leaf_batch` warps roll out from every selected child at once (1 = the reference's
        # one simulation per iteration); `iterations` stays the budget in SIMULATIONS
        self.leaf_batch = max(1, min(16, int(leaf_batch)))
        # One move at a time is what this agent is asked for, so by default the search is the round-synchronous one
        # that spreads ONE tree over the whole GPU (mcts.WideUCT, csrc/mcts_wide.cuh): `search_width` simulations are
        # planned per round (default: a quarter of the budget, 64..1024).  search_width=0 (or leaf_batch > 1) selects
        # the sequential one-warp / one-CTA search instead.
        self.search_width = search_width
        # root-parallel trees of the wide search: the simulation budget is split over `num_trees` bandits that are planned
        # side by side (one planner each) and share the GPU for their rollouts; their root statistics are summed.
        # Default: one tree per 2,500 simulations of the budget (rounded), at most 8.  (A round costs about the same whatever
        # its width since the flat rounds are dealt out at once, so extra planners no longer buy time; fewer, larger bandits
        # adapt better.  Round 2 began with one tree per 625 simulations, when a planner walked every simulation.)
        self.num_trees = num_trees
        self._thinks = 0
        self._wide_cache = None

    def select_action(self, board: Board, player: Player, legal_moves: List[Move]) -> Optional[Move]:
        return self.think(board, player, legal_moves, int(self.time_limit * 1000))["move"]

    @staticmethod
    def _move_dict(m: Move, visits: int, q: float) -> Dict[str, Any]:
        return {"piece_id": m.piece_id, "orientation": m.orientation, "anchor_row": m.anchor_row,
                "anchor_col": m.anchor_col, "visits": int(visits), "q_value": round(float(q), 4)}

    def think(self, board: Board, player: Player, legal_moves: List[Move], time_budget_ms: int) -> Dict[str, Any]:
        budget_s = max(time_budget_ms, 1) / 1000.0
        start = time.perf_counter()
        if not legal_moves:
            return {"move": None, "stats": {"timeBudgetMs": time_budget_ms, "timeSpentMs": 0, "nodesEvaluated": 0,
                                            "maxDepthReached": 0, "topMoves": []}}
        if len(legal_moves) == 1:
            m = legal_moves[0]
            return {"move": m, "stats": {"timeBudgetMs": time_budget_ms,
                                         "timeSpentMs": int((time.perf_counter() - start) * 1000),
                                         "nodesEvaluated": 1, "maxDepthReached": 1,
                                         "topMoves": [dict(self._move_dict(m, 1, 0.0), q_value=0.0)]}}
        # rank <-> move: ascending id.  The generator's own list says so itself; anything else is sorted here
        ids_known = getattr(legal_moves, "ascending_ids", None)
        ordered = legal_moves if ids_known is not None and len(ids_known) == len(legal_moves) else sorted(legal_moves, key=lambda m: m.to_id())
        wide = self.leaf_batch == 1 and (self.search_width is None or int(self.search_width) > 0)
        if wide:
            W = int(self.num_trees) if self.num_trees else max(1, min(8, (int(self.iterations) + 1250) // 2500))  # trees
            total = max(1, -(-int(self.iterations) // W))         # simulations per tree
            width = int(self.search_width) if self.search_width else max(64, min(1024, total // 4))
            key = (total, width, W, self.exploration_constant, self.rollout_plies, self.seed)
            if self._wide_cache is not None and self._wide_cache[0] == key:
                eng = self._wide_cache[1]                         # workspace and tables are kept between moves
                eng.p.tree_offset = self._thinks * W
            else:
                eng = WideUCT(total, width=width, n_trees=W, max_rollout_moves=self.rollout_plies, max_depth=1, ucb_variant=1,
                              exploration=self.exploration_constant, seed=self.seed, tree_offset=self._thinks * W, reward_mode=1)
                self._wide_cache = (key, eng)
        else:
            W = self.leaf_batch
            total = max(1, -(-int(self.iterations) // W))            # device iterations of W simulations each
            eng = DeviceUCT(total, n_trees=1, max_rollout_moves=self.rollout_plies, max_depth=1, ucb_variant=1,
                            exploration=self.exploration_constant, seed=self.seed, tree_offset=self._thinks,
                            reward_mode=1, leaf_batch=W)
        self._thinks += 1
        root = _root_state(board, player)
        n_root = len(ordered)
        done = 0
        trace: List[Dict[str, Any]] = []
        res = None
        interval = max(1, self.diagnostics_sample_interval // W)    # in device iterations
        next_sample = interval
        chunk = max(self.CHUNK, 4 * width) if wide else self.CHUNK
        while done < total and (done == 0 or time.perf_counter() - start < budget_s):
            n = min(chunk, total - done)
            if self.enable_diagnostics:
                n = min(n, next_sample - done)          # land exactly on the sampling points (:164-175)
            res = eng.search(root if done == 0 else None, n, resume=done > 0)
            done += n
            if not wide:                                  # (bk_mcts_wide_search returns when its rounds have finished)
                torch.cuda.current_stream().synchronize()
            if self.enable_diagnostics and done == next_sample:
                next_sample += interval
                if done < total:
                    v = res["visits"][:, :n_root].sum(dim=0).cpu().numpy() if wide else res["visits"][0, :n_root].cpu().numpy()
                    w = res["reward"][:, :n_root].sum(dim=0).cpu().numpy() if wide else res["reward"][0, :n_root].cpu().numpy()
                    b = int(np.argmax(v))
                    bm = ordered[b]
                    trace.append({"sim": done * W,
                                  "bestActionId": f"{bm.piece_id}-{bm.orientation}-{bm.anchor_row}-{bm.anchor_col}",
                                  "bestQMean": float(w[b] / v[b]) if v[b] > 0 else 0.0,
                                  "entropy": float(compute_policy_entropy([int(x) for x in v if x > 0]))})
        # one read-back: summed root statistics (over the trees of the wide search) and the device's legal-move count
        sel = slice(None) if wide else slice(0, 1)
        both = torch.cat([res["visits"][sel, :n_root].sum(dim=0, dtype=torch.int64), res["reward"][sel, :n_root].sum(dim=0),
                          res["counts"][:1].to(torch.int64)]).cpu().numpy()
        if int(both[-1]) != n_root:
            raise dev._lib.BlokusB200Error("legal_moves does not match the device's legal-move set for this board")
        visits, reward = both[:n_root], both[n_root:2 * n_root]
        # children were created from the highest rank down; "first child with max visits" = highest rank on ties
        ranks_desc = np.arange(n_root - 1, -1, -1)
        ranks_desc = ranks_desc[visits[ranks_desc] > 0]
        by_visits = ranks_desc[np.argsort(-visits[ranks_desc], kind="stable")]
        expanded = [int(r) for r in ranks_desc[::-1]]
        top_moves = [self._move_dict(ordered[r], visits[r], reward[r] / visits[r]) for r in by_visits[:10]]
        spent_ms = int((time.perf_counter() - start) * 1000)
        done *= W                                                  # from here on: simulations
        if done < 5:        # fast_mcts_agent.py:179-191: too few simulations -> quick heuristic pick
            pick = sorted(sorted(legal_moves, key=lambda m: m.piece_id, reverse=True)[:3],
                          key=lambda m: abs(m.anchor_row - 9.5) + abs(m.anchor_col - 9.5))[0]
            return {"move": pick, "stats": {"timeBudgetMs": time_budget_ms, "timeSpentMs": spent_ms,
                                            "nodesEvaluated": max(done, 1), "maxDepthReached": 2, "topMoves": top_moves}}
        best_rank = int(by_visits[0]) if len(by_visits) else 0
        entropy = compute_policy_entropy([int(visits[r]) for r in expanded]) if self.enable_diagnostics else 0.0
        n_exp = len(expanded)
        diagnostics = None
        if self.enable_diagnostics:
            diagnostics = {
                "version": "v1", "timeBudgetMs": int(time_budget_ms), "timeSpentMs": int(spent_ms),
                "simulations": max(done, 1),
                "simsPerSec": int(max(done, 1) / (spent_ms / 1000.0)) if spent_ms > 0 else 0,
                "rootLegalMoves": n_root, "rootChildrenExpanded": n_exp, "rootPolicy": top_moves,
                "policyEntropy": float(entropy), "maxDepthReached": 1 if n_exp else 0, "nodesExpanded": n_exp,
                "nodesByDepth": [{"depth": 0, "nodes": 1}] + ([{"depth": 1, "nodes": n_exp}] if n_exp else []),
                "bestMoveTrace": trace,
            }
        return {"move": ordered[best_rank],
                "stats": {"timeBudgetMs": time_budget_ms, "timeSpentMs": spent_ms, "nodesEvaluated": max(done, 1),
                          "maxDepthReached": (1 if n_exp else 0) if self.enable_diagnostics else 2,
                          "topMoves": top_moves, "diagnostics": diagnostics}}


class GameplayFastMCTSAgent:
    """``choose_move(board, player, legal_moves, time_budget_ms) -> (move, stats)`` (gameplay_fast_mcts.py:29-37)."""

    def __init__(self, *, iterations: int = 5000, exploration_constant: float = 1.414, seed: Optional[int] = None) -> None:
        self._agent = FastMCTSAgent(iterations=iterations, time_limit=1.0, exploration_constant=exploration_constant,
                                    seed=seed)

    def choose_move(self, board: Board, player: Player, legal_moves: List[Move],
                    time_budget_ms: int) -> Tuple[Optional[Move], Dict[str, Any]]:
        result = self._agent.think(board, player, legal_moves, time_budget_ms)
        return result.get("move"), dict(result.get("stats") or {})

"""Device UCT search (K5) behind the reference's search interfaces.

``DeviceUCT`` is the batched engine (n_trees root-parallel trees, one warp each, see csrc/mcts.cuh);
``ZobristHash`` / ``TranspositionTable`` follow ``mcts/zobrist.py``; ``MCTSAgent`` follows
``mcts/mcts_agent.py:194-341`` (constructor arguments, ``select_action``, ``stats``, ``reset``,
``set_seed``, ``get_action_info``).
"""

from __future__ import annotations

import ctypes as C
import time
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import _lib
from . import device as dev
from .board import Board, Player
from .move_generator import Move, get_shared_generator

MAX_ROOT_MOVES = 2048


# ---------------------------------------------------------------------------------------------------
# Zobrist keys and transposition table (mcts/zobrist.py)
# ---------------------------------------------------------------------------------------------------
class ZobristHash:
    """64-bit keys drawn exactly as the reference draws them: ``np.random.RandomState(seed)`` and one
    ``randint(0, 2**64, dtype=uint64)`` per entry in the order cells[20][20][5], turn[4], pieces[4][21]
    (mcts/zobrist.py:33-68).  Hashing runs on the device (csrc: zobrist_of)."""

    def __init__(self, board_size: int = 20, num_players: int = 4, seed: Optional[int] = None):
        if board_size != 20 or num_players != 4:
            raise ValueError("the device tables are built for the 20x20, 4-player game")
        self.board_size, self.num_players = board_size, num_players
        self.rng = np.random.RandomState(seed)
        n = board_size * board_size * (num_players + 1)
        flat = np.array([self.rng.randint(0, 2**64, dtype=np.uint64) for _ in range(n + num_players + num_players * 21)],
                        dtype=np.uint64)
        self.position_player_hashes = flat[:n].reshape(board_size, board_size, num_players + 1)
        self.player_turn_hashes = flat[n:n + num_players]
        self.piece_used_hashes = flat[n + num_players:].reshape(num_players, 21)
        self.table = flat
        self._dev: Optional[torch.Tensor] = None

    def device_table(self) -> torch.Tensor:
        if self._dev is None:
            self._dev = torch.from_numpy(self.table.view(np.int64).copy()).cuda()
        return self._dev

    def hash_board(self, board: Board) -> int:
        st = dev.to_device(np.atleast_1d(board.to_packed()))
        return int(dev.zobrist_hash(st, self.device_table()).cpu().numpy().view(np.uint64)[0])

    def hash_states(self, states: torch.Tensor) -> torch.Tensor:
        return dev.zobrist_hash(states, self.device_table())

    def hash_move(self, board: Board, move_hash: int, player: Player, piece_id: int) -> int:
        """Turn and spent-piece part of an incremental update (mcts/zobrist.py:101-128): plain table arithmetic."""
        cur = board.current_player.value - 1
        h = int(move_hash) ^ int(self.player_turn_hashes[cur]) ^ int(self.player_turn_hashes[(cur + 1) % self.num_players])
        return h ^ int(self.piece_used_hashes[player.value - 1, piece_id - 1])

    def hash_position_placement(self, row: int, col: int, player: Player) -> int:
        """mcts/zobrist.py:130-142."""
        return int(self.position_player_hashes[row, col, player.value])

    def get_hash_info(self) -> Dict[str, Any]:
        return {"board_size": self.board_size, "num_players": self.num_players, "hash_bits": 64,
                "total_positions": self.board_size * self.board_size, "total_pieces": 21}


class TranspositionTable:
    """key -> {"reward": r} with the reference's method names (mcts/zobrist.py:155-220), stored in the
    open-addressing device table the search kernel probes.  ``max_size`` rounds up to a power of two."""

    def __init__(self, max_size: int = 1000000):
        self.max_size = max_size
        self._t = dev.DeviceTranspositionTable(max(4, int(np.ceil(np.log2(max(16, 2 * max_size))))))
        self.access_count = 0
        self.hit_count = 0

    @property
    def device(self) -> dev.DeviceTranspositionTable:
        return self._t

    def get(self, hash_value: int) -> Optional[Dict]:
        self.access_count += 1
        k = torch.tensor([np.uint64(hash_value).astype(np.int64)], dtype=torch.int64, device="cuda")
        r, f = self._t.get(k)
        if int(f.item()):
            self.hit_count += 1
            return {"reward": float(r.item())}
        return None

    def put(self, hash_value: int, entry: Dict):
        self.evict_if_full()
        k = torch.tensor([np.uint64(hash_value).astype(np.int64)], dtype=torch.int64, device="cuda")
        self._t.put(k, torch.tensor([float(entry["reward"])], dtype=torch.float64, device="cuda"))

    def evict_if_full(self) -> None:
        """mcts/zobrist.py:196-200: a put into a full table first removes the max_size // 10 entries inserted first.
        The search kernel inserts without this check (the device table has >= 2 x max_size slots); DeviceUCT callers
        run it between searches."""
        if len(self) >= self.max_size:
            self._t.evict_oldest(self.max_size // 10)

    def search_counters(self) -> Dict[str, int]:
        """Hits / probes of the search kernels that used this table (counted on the device)."""
        st = self._t.stats()
        return {"hits": st["hits"], "probes": st["probes"]}

    def clear(self):
        self._t.clear()
        self.access_count = 0
        self.hit_count = 0

    def __len__(self):
        return self._t.stats()["size"]

    def get_stats(self) -> Dict[str, float]:
        st = self._t.stats()
        access, hits = self.access_count + st["probes"], self.hit_count + st["hits"]     # host calls + in-kernel probes
        return {"size": st["size"], "max_size": self.max_size, "access_count": access,
                "hit_count": hits, "hit_rate": hits / max(access, 1)}


# ---------------------------------------------------------------------------------------------------
# the device search engine
# ---------------------------------------------------------------------------------------------------
class DeviceUCT:
    """n_trees independent UCT trees over the same (or a few) roots; results per tree by root-move rank."""

    def __init__(self, iterations_total: int, n_trees: int = 1, max_rollout_moves: int = 50, max_depth: int = 0,
                 ucb_variant: int = 0, exploration: float = 1.414, seed: int = 0, tree_offset: int = 0,
                 reward_mode: int = 0, rollout_policy: int = 0, tt: Optional[dev.DeviceTranspositionTable] = None, zobrist: Optional[ZobristHash] = None,
                 max_root_moves: int = MAX_ROOT_MOVES, leaf_batch: int = 1, leaf_model: Optional[torch.Tensor] = None,
                 leaf_max_turns: int = 2500, progressive_bias_weight: float = 0.0, shaping_weight: float = 1.0,
                 shaping_gamma: float = 1.0, potential_mode: str = "prob", leaf_model_kind: int = 0, tt_tree_base: int = 0):
        if not torch.cuda.is_available():
            raise _lib.BlokusB200Error("no CUDA device: blokus_b200 has no CPU fallback")
        self.lib = _lib.init(torch.cuda.current_device())
        self.p = _lib.MctsParams(iterations=0, iterations_total=int(iterations_total),
                                 max_rollout_moves=int(max_rollout_moves), rollout_policy=int(rollout_policy), max_depth=int(max_depth),
                                 ucb_variant=int(ucb_variant), resume=0, reward_mode=int(reward_mode), leaf_batch=int(leaf_batch),
                                 tt_tree_base=int(tt_tree_base), leaf_max_turns=int(leaf_max_turns), leaf_model_kind=int(leaf_model_kind),
                                 progressive_bias_weight=float(progressive_bias_weight) if int(reward_mode) >= 3 else 0.0,
                                 shaping_weight=float(shaping_weight), shaping_gamma=float(shaping_gamma),
                                 potential_mode=1 if potential_mode == "logit" else 0, round_width=0,
                                 exploration=float(exploration),
                                 seed=int(seed) & (2**64 - 1), tree_offset=int(tree_offset))
        self.n_trees = int(n_trees)
        # reward_mode 3 / 4: learned leaf evaluation / rollouts with potential shaping, both with a folded pairwise
        # logistic model (float64[35] on the device)
        self.leaf_model = None
        if int(reward_mode) >= 3:
            if leaf_model is None or (int(leaf_model_kind) == 0 and leaf_model.numel() != 35):
                raise ValueError("reward_mode 3 / 4 needs leaf_model: 34 weights + bias, or a flattened GBT blob (kind 1)")
            self.leaf_model = leaf_model.to(device="cuda", dtype=torch.float64).contiguous()
            self.p.leaf_model = self.leaf_model.data_ptr()
        self.root_frontiers = None
        self.leaf_batch = max(1, int(leaf_batch))       # simulations per iteration (warps per tree)
        if self.leaf_batch > 16:
            raise ValueError("leaf_batch must be in [1, 16]")
        self.max_root_moves = int(max_root_moves)
        self.tt, self.zobrist = tt, zobrist
        if tt is not None and zobrist is None:
            raise ValueError("a transposition table needs a ZobristHash")
        per_tree = int(self.lib.bk_mcts_workspace_bytes(C.byref(self.p)))
        self.ws_bytes = per_tree * self.n_trees
        d = torch.device("cuda", torch.cuda.current_device())
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=d)
        m = self.max_root_moves
        self.visits = torch.zeros((self.n_trees, m), dtype=torch.int32, device=d)
        self.reward = torch.zeros((self.n_trees, m), dtype=torch.int64, device=d)
        self.moves = torch.zeros((self.n_trees, m), dtype=torch.uint16, device=d)
        self.counts = torch.zeros(self.n_trees, dtype=torch.int32, device=d)
        self.status = torch.zeros(self.n_trees, dtype=torch.int32, device=d)
        self.iterations_done = 0
        self._roots = None

    def search(self, roots: Optional[torch.Tensor], iterations: int, resume: bool = False,
               root_frontiers: Optional[torch.Tensor] = None, active_trees: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """Run `iterations` more iterations on every tree (fresh trees from `roots` unless resume).  root_frontiers
        (int64[n_roots, 4, 7], learned leaf evaluation only): the roots' incremental frontier sets; None = rebuilt.
        active_trees (<= n_trees) searches only the first so many trees (a batch whose size changes from call to call);
        rows beyond it in the returned tensors are stale."""
        if not resume:
            if roots is None:
                raise ValueError("roots required for a fresh search")
            self._roots = roots.contiguous()
            self.iterations_done = 0
            self.root_frontiers = None
            if root_frontiers is not None:
                if tuple(root_frontiers.shape) != (self._roots.shape[0], 4, 7) or root_frontiers.dtype != torch.int64:
                    raise ValueError("root_frontiers must be int64[n_roots, 4, 7]")
                self.root_frontiers = root_frontiers.to(self._roots.device).contiguous()
            self.p.root_frontiers = self.root_frontiers.data_ptr() if self.root_frontiers is not None else None
        if self.iterations_done + iterations > self.p.iterations_total:
            raise ValueError("search exceeds iterations_total")
        self.p.iterations = int(iterations)
        self.p.resume = 1 if resume else 0
        n_roots = self._roots.shape[0]
        if not resume:
            self._active = self.n_trees if active_trees is None else int(active_trees)
            if not 1 <= self._active <= self.n_trees:
                raise ValueError("active_trees must be in [1, n_trees]")
        check = _lib.check
        check(self.lib.bk_mcts_search(
            self._roots.data_ptr(), n_roots, self._active, C.byref(self.p),
            self.tt.handle if self.tt is not None else None,
            self.zobrist.device_table().data_ptr() if self.zobrist is not None else None,
            self.workspace.data_ptr(), self.ws_bytes, self.max_root_moves,
            self.visits.data_ptr(), self.reward.data_ptr(), self.counts.data_ptr(), self.moves.data_ptr(),
            self.status.data_ptr(), dev._stream()))
        dev._count()
        self.iterations_done += iterations
        reward = self.reward.view(torch.float64) if self.p.reward_mode >= 3 else self.reward
        return {"visits": self.visits, "reward": reward, "moves": self.moves, "counts": self.counts,
                "status": self.status}


class WideUCT:
    """ONE tree (per root) spread over the whole GPU: the round-synchronous search of csrc/mcts_wide.cuh.

    ``DeviceUCT`` runs the reference's sequential iteration with one warp per tree -- the right shape for thousands of
    trees, 1 of 148 SMs for a single move.  ``WideUCT`` plans ``width`` simulations per round with UCB1 walks under
    virtual loss and runs them on the whole grid; 1,000 simulations take a few rounds instead of 1,000 sequential
    rollouts.  Deterministic in (seed, tree_offset, width, sequence of search calls); integer reward modes only
    (0 = MCTSAgent's leaf-player gain, 1 = root-player gain of the flat bandit, 2 = constant leaf), no transposition
    table.  ``search`` returns when the search has finished (the C call synchronises the stream)."""

    def __init__(self, sims_total: int, width: int = 256, n_trees: int = 1, max_rollout_moves: int = 50, max_depth: int = 0,
                 ucb_variant: int = 0, exploration: float = 1.414, seed: int = 0, tree_offset: int = 0, reward_mode: int = 0,
                 rollout_policy: int = 0, max_root_moves: int = MAX_ROOT_MOVES):
        if not torch.cuda.is_available():
            raise _lib.BlokusB200Error("no CUDA device: blokus_b200 has no CPU fallback")
        if int(reward_mode) not in (0, 1, 2):
            raise ValueError("the wide search backs rewards up with integer atomics: reward_mode must be 0, 1 or 2")
        self.lib = _lib.init(torch.cuda.current_device())
        self.p = _lib.MctsParams(iterations=0, iterations_total=int(sims_total), max_rollout_moves=int(max_rollout_moves),
                                 rollout_policy=int(rollout_policy), max_depth=int(max_depth), ucb_variant=int(ucb_variant),
                                 resume=0, reward_mode=int(reward_mode), leaf_batch=1, tt_tree_base=0,
                                 exploration=float(exploration), seed=int(seed) & (2**64 - 1), tree_offset=int(tree_offset),
                                 round_width=int(width))
        self.n_trees, self.width, self.max_root_moves = int(n_trees), int(width), int(max_root_moves)
        self.ws_bytes = int(self.lib.bk_mcts_wide_workspace_bytes(C.byref(self.p))) * self.n_trees
        d = torch.device("cuda", torch.cuda.current_device())
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=d)
        m = self.max_root_moves
        self.visits = torch.zeros((self.n_trees, m), dtype=torch.int32, device=d)
        self.reward = torch.zeros((self.n_trees, m), dtype=torch.int64, device=d)
        self.moves = torch.zeros((self.n_trees, m), dtype=torch.uint16, device=d)
        self.counts = torch.zeros(self.n_trees, dtype=torch.int32, device=d)
        self.status = torch.zeros(self.n_trees, dtype=torch.int32, device=d)
        self.done = torch.zeros(self.n_trees, dtype=torch.int32, device=d)
        self.sims_done = 0
        self._roots = None

    def search(self, roots: Optional[torch.Tensor], sims: int, resume: bool = False) -> Dict[str, torch.Tensor]:
        """Add `sims` simulations to every tree (fresh trees from `roots` unless resume); tree k searches roots[k % n]."""
        if not resume:
            if roots is None:
                raise ValueError("roots required for a fresh search")
            self._roots = roots.contiguous()
            self.sims_done = 0
        if self.sims_done + sims > self.p.iterations_total:
            raise ValueError("search exceeds sims_total")
        self.p.iterations = int(sims)
        self.p.resume = 1 if resume else 0
        _lib.check(self.lib.bk_mcts_wide_search(
            self._roots.data_ptr(), self._roots.shape[0], self.n_trees, C.byref(self.p), self.workspace.data_ptr(), self.ws_bytes,
            self.max_root_moves, self.visits.data_ptr(), self.reward.data_ptr(), self.counts.data_ptr(), self.moves.data_ptr(),
            self.status.data_ptr(), self.done.data_ptr(), dev._stream()))
        rounds = -(-int(sims) // self.width) + 1
        dev._count(2 * rounds + 1)
        self.sims_done += int(sims)
        return {"visits": self.visits, "reward": self.reward, "moves": self.moves, "counts": self.counts,
                "status": self.status, "sims_done": self.done}


def _root_state(board: Board, player: Player) -> torch.Tensor:
    s = board.to_packed()
    s["to_move"] = player.value - 1          # MCTSNode(board, player): the root player moves (mcts_agent.py:325)
    return dev.to_device(np.atleast_1d(s))


def _best_by_visits_creation_order(visits: np.ndarray, n_root: int) -> int:
    """Rank of the reference's get_best_move(): first child with max visits in creation order; children are
    created from the last legal move to the first (expand() pops), so ties go to the highest rank."""
    best, best_v = -1, -1
    for rank in range(n_root - 1, -1, -1):
        if visits[rank] > best_v:
            best, best_v = rank, int(visits[rank])
    return best


# ---------------------------------------------------------------------------------------------------
# MCTSAgent (mcts/mcts_agent.py:194-632)
# ---------------------------------------------------------------------------------------------------
class MCTSAgent:
    """UCT agent; every simulation (select/expand/rollout/backprop) runs in one device kernel.

    Same constructor arguments as the reference.  Differences, all documented in DESIGN.md: rollout moves are
    drawn with Philox on the device (``rollout_agent`` objects are accepted for interface compatibility but only
    their kind -- heuristic softmax, the default, or uniform random -- matters); ``num_trees`` > 1 adds root-parallel
    trees whose root statistics are summed.

    Learned-evaluator options (mcts_agent.py:427-466, 486-551): with the reference's ``dummy_model.json`` artifact --
    the only one its recorded arena runs use -- every probability is 0.5, so leaf evaluation is the kernel's constant
    leaf value (reward_mode 2), the progressive bias term is exactly 0 and the shaping term is exactly 0 for gamma = 1
    or potential_mode "logit"; those combinations run on the device.  A fitted ``pairwise_logreg`` artifact whose
    pipeline is [StandardScaler,] LogisticRegression also runs in the tree: leaf_evaluation_enabled evaluates the 34
    snapshot features of every new leaf in the kernel and feeds the folded 34 weights + bias (reward_mode 3), and
    progressive_bias_enabled adds the selection term w * prior / (1 + n) with prior = V(child, parent player) -
    V(parent, parent player) from the same feature block; potential_shaping_enabled (without leaf evaluation) keeps
    the rollouts and adds w * (gamma * Phi(end) - Phi(leaf)) to truncated ones (reward_mode 4).  ``pairwise_gbt_phase``
    artifacts (GradientBoostingClassifier per game phase) run the same way with their trees flattened for the kernel.
    Artifacts of any other shape are served out of tree by ``winprob.LearnedWinProbabilityEvaluator``.
    """

    def __init__(self, iterations: int = 1000, time_limit: Optional[float] = None, exploration_constant: float = 1.414,
                 rollout_agent=None, use_transposition_table: bool = True, seed: Optional[int] = None,
                 learned_model_path: Optional[str] = None, leaf_evaluation_enabled: bool = False,
                 progressive_bias_enabled: bool = False, progressive_bias_weight: float = 0.25,
                 potential_shaping_enabled: bool = False, potential_shaping_gamma: float = 1.0,
                 potential_shaping_weight: float = 1.0, potential_mode: str = "prob", max_rollout_moves: int = 50,
                 num_trees: int = 1, leaf_batch: int = 1, search_width: int = 0):
        if potential_mode not in {"prob", "logit"}:
            raise ValueError("potential_mode must be either 'prob' or 'logit'.")
        if int(max_rollout_moves) <= 0:
            raise ValueError("max_rollout_moves must be > 0.")
        self.leaf_evaluation_enabled = bool(leaf_evaluation_enabled)
        self.progressive_bias_enabled = bool(progressive_bias_enabled)
        self.progressive_bias_weight = float(progressive_bias_weight)
        self.potential_shaping_enabled = bool(potential_shaping_enabled)
        self.potential_shaping_gamma = float(potential_shaping_gamma)
        self.potential_shaping_weight = float(potential_shaping_weight)
        self.potential_mode = potential_mode
        self.learned_model_path = learned_model_path
        self._requires_learned_evaluator = (self.leaf_evaluation_enabled or self.progressive_bias_enabled
                                            or self.potential_shaping_enabled)
        if self._requires_learned_evaluator and not learned_model_path:
            raise ValueError("learned_model_path is required when learned evaluation, progressive bias, "
                             "or potential shaping is enabled.")
        self.learned_evaluator = None
        if learned_model_path:
            from .winprob import LearnedWinProbabilityEvaluator

            self.learned_evaluator = LearnedWinProbabilityEvaluator(learned_model_path, potential_mode=potential_mode)
            real = not self.learned_evaluator._is_dummy
            if real and self.progressive_bias_enabled and not (self.leaf_evaluation_enabled or self.potential_shaping_enabled):
                raise NotImplementedError("progressive bias with a fitted model runs in the learned kernels: enable "
                                          "leaf evaluation or potential shaping with it")
            if real and self._requires_learned_evaluator and getattr(self.learned_evaluator, "device_model", None) is None:
                raise NotImplementedError("the learned options run in the kernel for pairwise_logreg artifacts whose "
                                          "pipeline is [StandardScaler,] LogisticRegression and for pairwise_gbt_phase "
                                          "artifacts of scikit-learn GradientBoostingClassifier models; this artifact has "
                                          "another shape -- use winprob.LearnedWinProbabilityEvaluator out of tree")
            if (not real and self.potential_shaping_enabled and potential_mode == "prob"
                    and self.potential_shaping_gamma != 1.0 and self.potential_shaping_weight != 0.0):
                raise NotImplementedError("potential shaping with gamma != 1 adds a non-zero constant to truncated "
                                          "rollouts; only the zero-term configurations run on the device")
        self.iterations = iterations
        self.time_limit = time_limit
        self.exploration_constant = exploration_constant
        self.use_transposition_table = use_transposition_table
        self.max_rollout_moves = int(max_rollout_moves)
        self.num_trees = int(num_trees)
        # leaf-batched rollouts: every iteration rolls out `leaf_batch` times from the new leaf (1 = the reference);
        # with the transposition table a batch caches the sum of its rewards
        self.leaf_batch = max(1, int(leaf_batch))
        # search_width > 0: ONE tree over the whole GPU (WideUCT: `search_width` simulations planned per round, then run
        # on the whole grid) instead of the reference's sequential iteration -- a different, deterministic schedule for
        # callers that need one move fast; rollouts or the dummy leaf value only, the transposition table is not used
        self.search_width = max(0, int(search_width))
        self.seed = 0 if seed is None else int(seed)
        self.move_generator = get_shared_generator()
        self.zobrist_hash = ZobristHash(seed=seed)
        # rollout policy: HeuristicAgent by default, as in the reference (mcts_agent.py:275-276); an agent whose
        # get_action_info() says "random" selects the uniform policy.  Either way the draw itself is Philox.
        self.rollout_agent = rollout_agent
        kind = "heuristic"
        if rollout_agent is not None:
            try:
                kind = rollout_agent.get_action_info().get("type", "heuristic")
            except Exception:
                kind = "heuristic"
        self.rollout_policy = dev.POLICY_UNIFORM if kind == "random" else dev.POLICY_HEURISTIC
        self.transposition_table = TranspositionTable() if use_transposition_table else None
        self._searches = 0
        self.reset_stats()

    def reset_stats(self):
        self.stats = {"iterations_run": 0, "time_elapsed": 0.0, "transposition_hits": 0, "rollout_rewards": [],
                      "leaf_eval_calls": 0, "progressive_bias_updates": 0, "potential_shaping_terms": [],
                      "evaluator_errors": 0}

    def _engine(self, total: int) -> DeviceUCT:
        tt = self.transposition_table.device if self.transposition_table is not None else None
        # every search of this agent gets its own block of Philox streams; the transposition-table namespaces stay
        # 0 .. num_trees-1 over all its searches, so a later search finds what an earlier one cached
        # (mcts_agent.py:337-339: the reference keeps one table per agent and clears it only past 500 000 entries)
        offset = self._searches * max(self.num_trees, 1)
        self._searches += 1
        if self.transposition_table is not None:
            self.transposition_table.evict_if_full()
        # the engine (workspace, result buffers) is kept across the moves of the agent; only the stream block changes
        key = (int(total), self.seed, id(self.zobrist_hash), self.exploration_constant, self.max_rollout_moves)
        cached = getattr(self, "_engine_cache", None)
        if cached is not None and cached[0] == key:
            cached[1].p.tree_offset = int(offset)
            return cached[1]
        self._engine_cache = None
        mode, model = 0, None
        real = self.learned_evaluator is not None and not self.learned_evaluator._is_dummy
        if self.leaf_evaluation_enabled:
            if self.learned_evaluator._is_dummy:
                mode = 2
            else:
                mode, model = 3, torch.from_numpy(self.learned_evaluator.device_model)
        elif self.potential_shaping_enabled and real:            # rollouts + shaping term on truncated ones
            mode, model = 4, torch.from_numpy(self.learned_evaluator.device_model)
        eng = DeviceUCT(total, n_trees=self.num_trees, max_rollout_moves=self.max_rollout_moves,
                         exploration=self.exploration_constant, seed=self.seed, tree_offset=offset,
                         reward_mode=mode, leaf_model=model, leaf_max_turns=getattr(self.learned_evaluator, "max_turns", 2500),
                         progressive_bias_weight=self.progressive_bias_weight if self.progressive_bias_enabled else 0.0,
                         leaf_model_kind=getattr(self.learned_evaluator, "device_model_kind", 0),
                         shaping_weight=self.potential_shaping_weight, shaping_gamma=self.potential_shaping_gamma,
                         potential_mode=self.potential_mode,
                         leaf_batch=self.leaf_batch, tt_tree_base=0,
                         rollout_policy=self.rollout_policy, tt=tt,
                         zobrist=self.zobrist_hash if tt is not None else None)
        self._engine_cache = (key, eng)
        return eng

    def search_root(self, board: Board, player: Player) -> Dict[str, np.ndarray]:
        """Root statistics (visits, reward sums) by ascending legal-move rank, summed over the trees."""
        start = time.time()
        root = _root_state(board, player)
        hits0 = self.transposition_table.search_counters()["hits"] if self.transposition_table is not None else 0
        frontiers = None
        if self._requires_learned_evaluator and self.learned_evaluator is not None and not self.learned_evaluator._is_dummy:
            frontiers = torch.from_numpy(board.frontier_words()[None].copy())     # the board's incremental sets
        if self.time_limit:
            # capacity for the time budget: one warp runs about 12-15 k simulations/s (profiles/r01_agent_latency.json);
            # the workspace grows with the capacity (child pool up to 2048 slots per iteration), so the bound is tight
            chunk = 64
            cap = min(60000, max(4 * chunk, int(self.time_limit * 16000)))
            eng = self._engine(cap)
            done = 0
            res = eng.search(root, chunk, root_frontiers=frontiers)
            done += chunk
            torch.cuda.synchronize()
            while time.time() - start < self.time_limit and done + chunk <= cap:
                res = eng.search(None, chunk, resume=True)
                done += chunk
                torch.cuda.synchronize()
        elif self.search_width > 0:
            if self._requires_learned_evaluator and not self.learned_evaluator._is_dummy:
                raise NotImplementedError("search_width > 0 runs rollouts or the dummy leaf value; fitted models run in the "
                                          "sequential kernel (search_width=0)")
            total = max(1, int(self.iterations))
            key = ("wide", total, self.search_width, self.seed, self.exploration_constant, self.max_rollout_moves)
            cached = getattr(self, "_wide_cache", None)
            if cached is not None and cached[0] == key:
                eng = cached[1]
                eng.p.tree_offset = self._searches
            else:
                eng = WideUCT(total, width=self.search_width, max_rollout_moves=self.max_rollout_moves,
                              exploration=self.exploration_constant, seed=self.seed, tree_offset=self._searches,
                              reward_mode=2 if self.leaf_evaluation_enabled else 0, rollout_policy=self.rollout_policy)
                self._wide_cache = (key, eng)
            self._searches += 1
            res = eng.search(root, total)
            done = total
        else:
            eng = self._engine(max(1, int(self.iterations)))
            res = eng.search(root, int(self.iterations), root_frontiers=frontiers)
            done = int(self.iterations)
        if int(res["status"].max().item()):
            raise _lib.BlokusB200Error("MCTS workspace exhausted")
        n_root = int(res["counts"][0].item())
        visits = res["visits"][:, :n_root].sum(dim=0).cpu().numpy()
        reward = res["reward"][:, :n_root].sum(dim=0).cpu().numpy()
        self.stats["iterations_run"] = done
        self.stats["time_elapsed"] = time.time() - start
        if self.transposition_table is not None:         # mcts_agent.py:423, counted by the kernel
            self.stats["transposition_hits"] += self.transposition_table.search_counters()["hits"] - hits0
        if self.leaf_evaluation_enabled:
            if self.learned_evaluator._is_dummy:
                reward = reward * 0.5                        # halves -> the dummy evaluator's probabilities
            self.stats["leaf_eval_calls"] = done
        if self.progressive_bias_enabled:
            self.stats["progressive_bias_updates"] = done    # one (zero-valued) prior per expansion
        return {"visits": visits, "reward": reward, "n_root": n_root}

    def select_action(self, board: Board, player: Player, legal_moves: List[Move]) -> Optional[Move]:
        if not legal_moves:
            return None
        if len(legal_moves) == 1:
            return legal_moves[0]
        r = self.search_root(board, player)
        if self.num_trees == 1:
            rank = _best_by_visits_creation_order(r["visits"], r["n_root"])
        else:
            rank = int(np.argmax(r["visits"]))          # root-parallel: max summed visits, lowest id on ties
        ids = sorted(m.to_id() for m in legal_moves)            # rank <-> move: ascending canonical id
        if len(ids) != r["n_root"]:
            ids = self.move_generator.get_legal_move_ids(board, player)
        best = Move.from_id(int(ids[rank]))
        if self.transposition_table is not None and len(self.transposition_table) > 500000:
            self.transposition_table.clear()             # mcts_agent.py:337-339
        return best

    def get_action_info(self) -> Dict[str, Any]:
        info = {"name": "MCTSAgent", "type": "mcts",
                "description": "Monte Carlo Tree Search with UCT, device rollouts (B200)",
                "parameters": {"iterations": self.iterations, "time_limit": self.time_limit,
                               "exploration_constant": self.exploration_constant,
                               "use_transposition_table": self.use_transposition_table,
                               "max_rollout_moves": self.max_rollout_moves, "num_trees": self.num_trees},
                "stats": dict(self.stats)}
        if self.transposition_table is not None:
            info["transposition_stats"] = self.transposition_table.get_stats()
        return info

    def reset(self):
        self.reset_stats()
        if self.transposition_table is not None:
            self.transposition_table.clear()

    def set_seed(self, seed: int):
        self.seed = int(seed)
        self.zobrist_hash = ZobristHash(seed=seed)
        self._searches = 0


# ---------------------------------------------------------------------------------------------------
def smoke() -> None:
    """Tiny search used by __graft_entry__.smoke(): visits add up and the root count is right."""
    from .state import empty_states

    eng = DeviceUCT(64, n_trees=2, seed=3)
    res = eng.search(dev.to_device(empty_states(1)), 64)
    torch.cuda.synchronize()
    assert int(res["counts"][0]) == 58
    assert res["visits"].sum(dim=1).tolist() == [64, 64]


def bench_extras(n_trees: int = 1184, iterations: int = 256) -> Dict[str, float]:
    """MCTS simulations/s for bench.py: root-parallel trees from the empty board, 50-ply rollouts."""
    from .state import empty_states

    eng = DeviceUCT(iterations, n_trees=n_trees, seed=42)
    roots = dev.to_device(empty_states(1))
    eng.search(roots, iterations)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.search(roots, iterations)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"mcts_sims_per_sec": n_trees * iterations / ms * 1e3, "mcts_trees": n_trees,
            "mcts_iterations_per_tree": iterations, "mcts_rollout_plies": 50}

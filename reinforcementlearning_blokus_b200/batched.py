"""``BatchedBlokus``: N independent games resident in HBM, stepped together.

This is the shape the hot path wants on a GPU ("batched APIs must keep state on
device"): the state array never leaves the device; legal moves, steps, game-over masks and scores are
one kernel launch each for all N games.  Semantics per game are those of ``BlokusGame`` + the caller-side
pass of the reference's arena loop (analytics/tournament/arena_runner.py:653-695).
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import device as dev
from .state import empty_states

PASS = 0xFFFF


class BatchedBlokus:
    def __init__(self, n: int, states: Optional[torch.Tensor] = None):
        self.n = int(n)
        self.states = states.clone() if states is not None else dev.to_device(empty_states(self.n))
        if self.states.shape[0] != self.n:
            raise ValueError("states must have n rows")

    # ---- queries -------------------------------------------------------------------------------------
    def legal_moves(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """CSR (counts, offsets, ids) of the side to move of every game, ids ascending."""
        return dev.movegen(self.states)

    def legal_move_counts(self) -> torch.Tensor:
        return dev.movegen_count(self.states)

    def alive_mask(self) -> torch.Tensor:
        """uint8[N]: bit p set iff player p can move; a game is over when its mask is 0 (engine/game.py:182-214)."""
        return dev.has_moves(self.states)

    def scores(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return dev.scores(self.states)

    def to_move(self) -> torch.Tensor:
        return self.states[:, 240].clone()

    # ---- transitions ---------------------------------------------------------------------------------
    def step(self, move_ids: torch.Tensor, want_alive: bool = False):
        """make_move (or pass for 0xFFFF) for every game; returns (ok, alive) like device.step."""
        if move_ids.dtype == torch.int16:
            move_ids = move_ids.view(torch.uint16)
        return dev.step(self.states, move_ids.contiguous(), want_alive=want_alive)

    def step_uniform_random(self, seed: int, turn: int, game_offset: int = 0) -> torch.Tensor:
        """One turn of uniform-random play for every game (pass when no move): the single-turn slice of the
        playout kernel, drawing with the same Philox stream (seed, game, turn)."""
        r = dev.playout(self.n, seed=seed, game_offset=game_offset, start_states=self.states, max_turns=turn + 1,
                        want_trace=True, want_final=True, first_turn=turn)
        self.states = r["final_states"]
        return r["trace"][:, turn]

    def playout(self, seed: int, game_offset: int = 0, **kw) -> Dict[str, torch.Tensor]:
        """Finish every game with uniform-random play from the current states (states are not modified)."""
        return dev.playout(self.n, seed=seed, game_offset=game_offset, start_states=self.states, **kw)


def mobility_after_moves(state, player: int, mover: int, move_ids) -> "torch.Tensor":
    """Next-turn mobility of `player` after each of `mover`'s candidate moves, as one batch (row N3 of SURVEY.md
    section 8f): the inner loop of ``simulate_mobility_stability`` (engine/telemetry.py:17-77) -- copy the board, let
    an opponent place one sampled move, count the player's legal moves -- for all samples at once:
    one ``bk_step`` + one ``bk_movegen_count``.  Returns int32[len(move_ids)] (-1 where the move was illegal)."""
    ids = torch.as_tensor(np.asarray(move_ids, dtype=np.int64).astype(np.uint16).view(np.int16)).view(torch.uint16).cuda()
    k = ids.numel()
    base = dev.to_device(np.atleast_1d(state))
    states = base.repeat(k, 1).contiguous()
    movers = torch.full((k,), int(mover), dtype=torch.uint8, device=states.device)
    ok, _ = dev.step(states, ids, players=movers)
    counts = dev.movegen_count(states, torch.full((k,), int(player), dtype=torch.uint8, device=states.device))
    return torch.where(ok.bool(), counts, torch.full_like(counts, -1))


def selfplay_mcts(n_games: int, iterations: int, *, seed: int = 0, trees_per_game: int = 1, leaf_batch: int = 1,
                  rollout_policy: int = dev.POLICY_UNIFORM, max_rollout_moves: int = 50, exploration: float = 1.414,
                  game_offset: int = 0, max_turns: int = 96, want_policy: bool = False) -> Dict[str, torch.Tensor]:
    """Self-play where EVERY move of every game is chosen by UCT search -- the league's data generator
    (configs/stage3_selfplay_gpu.yaml feeds such games to training) as device batches: per turn one launch searches the
    current position of all n_games games (`trees_per_game` root-parallel trees per game, laid out game-major, their
    root statistics summed; tree j of global game G on turn t always draws from Philox stream id
    ((G + t * 2^24) * trees_per_game + j), whatever the batch or shard size), the most visited move is played (ties: the reference's rule for one tree,
    first child in creation order = highest rank), and games whose mover has no move pass
    (analytics/tournament/arena_runner.py:653-695).

    Returns trace uint16[N, T] (0xFFFF = pass), game_scores / board_scores int32[N, 4], turns int32[N] and, with
    want_policy, a list of (visits int32[N, M], move_ids uint16[N, M]) per turn -- visit-count policy targets."""
    from .mcts import MAX_ROOT_MOVES, DeviceUCT

    games = BatchedBlokus(n_games)
    n_trees = n_games * int(trees_per_game)
    eng = DeviceUCT(iterations, n_trees=n_trees, max_rollout_moves=max_rollout_moves, exploration=exploration, seed=seed,
                    rollout_policy=rollout_policy, leaf_batch=leaf_batch)
    trace = torch.full((n_games, max_turns), PASS, dtype=torch.int32, device=games.states.device)
    turns = torch.zeros(n_games, dtype=torch.int32, device=games.states.device)
    policy = []
    rank_weight = torch.arange(MAX_ROOT_MOVES, device=games.states.device, dtype=torch.int64)
    for t in range(max_turns):
        alive = games.alive_mask()
        running = alive != 0
        if not bool(running.any()):
            break
        # every (game, turn) pair gets its own block of Philox streams; trees are game-major (tree k searches root k)
        tpg = int(trees_per_game)
        eng.p.tree_offset = (int(game_offset) + t * (1 << 24)) * tpg
        roots = games.states if tpg == 1 else games.states.repeat_interleave(tpg, dim=0).contiguous()
        res = eng.search(roots, iterations)
        visits = res["visits"].view(n_games, tpg, -1).sum(dim=1)
        # most visited root move, ties to the highest rank; unexpanded ranks have 0 visits
        key = visits.to(torch.int64) * MAX_ROOT_MOVES + rank_weight[None, : visits.shape[1]]
        best = key.argmax(dim=1)
        moves_g = res["moves"].view(n_games, tpg, -1)[:, 0].contiguous()       # the move ids by rank are the same in every tree
        counts_g = res["counts"].view(n_games, tpg)[:, 0]
        ids = moves_g.view(torch.int16).gather(1, best[:, None]).view(torch.uint16)[:, 0]
        has_move = (counts_g > 0) & running
        move = torch.where(has_move, ids.to(torch.int32), torch.full_like(ids, PASS, dtype=torch.int32))
        move16 = move.to(torch.int16).view(torch.uint16).contiguous()      # 0..36399 and 0xFFFF survive the round trip
        before = games.states.clone()
        ok, _ = games.step(move16)
        games.states = torch.where(running[:, None], games.states, before)  # a finished game is not touched any more
        trace[:, t] = torch.where(running, move, trace[:, t])
        turns += running.to(torch.int32)
        if want_policy:
            policy.append((visits.clone(), moves_g.clone()))
    gs, bs = games.scores()
    out = {"trace": trace.to(torch.int16).view(torch.uint16), "game_scores": gs, "board_scores": bs, "turns": turns,
           "final_states": games.states}
    if want_policy:
        out["policy"] = policy
    return out

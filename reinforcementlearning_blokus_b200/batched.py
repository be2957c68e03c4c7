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

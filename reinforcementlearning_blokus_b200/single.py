"""Single-board calls for the drop-in (object) API: one 256-byte state up, one kernel, a few bytes back.

These calls are latency-bound by construction (SURVEY.md section 7 "host<->device friction"); the
throughput path is ``device.py`` / ``batched.py``.  Staging buffers are pinned and reused; a lock makes
the helpers callable from worker threads (the reference's web API runs agents in an executor thread,
webapi/app.py:330-333).
"""

from __future__ import annotations

import threading
from typing import List, Tuple

import numpy as np
import torch

from . import device as dev
from .state import STATE_BYTES

MAX_MOVES = 4096  # slots for one state's move list; an empty-board first move has 58, mid-game peaks are < 1000

_lock = threading.Lock()
_ctx = None


class _Ctx:
    def __init__(self):
        if not torch.cuda.is_available():
            raise dev._lib.BlokusB200Error("no CUDA device: blokus_b200 has no CPU fallback")
        d = torch.device("cuda", torch.cuda.current_device())
        self.lib = dev._lib_for(None)
        self.h_state = torch.empty((1, STATE_BYTES), dtype=torch.uint8).pin_memory()
        self.d_state = torch.empty((1, STATE_BYTES), dtype=torch.uint8, device=d)
        self.d_player = torch.empty(1, dtype=torch.uint8, device=d)
        self.h_player = torch.empty(1, dtype=torch.uint8).pin_memory()
        # move list and its length in ONE device buffer (ids, then the int32 count) so that one copy brings both back
        self.d_list = torch.empty(MAX_MOVES * 2 + 8, dtype=torch.uint8, device=d)
        self.h_list = torch.empty(MAX_MOVES * 2 + 8, dtype=torch.uint8).pin_memory()
        self.d_ids = self.d_list[: MAX_MOVES * 2].view(torch.uint16).view(1, MAX_MOVES)
        self.d_counts = self.d_list[MAX_MOVES * 2: MAX_MOVES * 2 + 4].view(torch.int32)
        self.h_ids_np = self.h_list.numpy()[: MAX_MOVES * 2].view(np.uint16)
        self.h_count_np = self.h_list.numpy()[MAX_MOVES * 2: MAX_MOVES * 2 + 4].view(np.int32)
        self.h_state_np = self.h_state.numpy()
        self.h_mask = torch.empty((1, 7), dtype=torch.int64).pin_memory()
        self.d_mask = torch.empty((1, 7), dtype=torch.int64, device=d)
        self.h_move = torch.empty(1, dtype=torch.uint16).pin_memory()
        self.d_move = torch.empty(1, dtype=torch.uint16, device=d)

    def upload(self, state_np, player: int) -> None:
        self.h_state_np[0] = np.frombuffer(np.asarray(state_np).tobytes(), dtype=np.uint8)
        self.h_player[0] = player
        self.d_state.copy_(self.h_state, non_blocking=True)
        self.d_player.copy_(self.h_player, non_blocking=True)

    def upload_as_mover(self, state_np, player: int) -> None:
        """One copy: the player rides in the state's side-to-move byte (the kernels default to it)."""
        self.h_state_np[0] = np.frombuffer(np.asarray(state_np).tobytes(), dtype=np.uint8)
        self.h_state_np[0, 240] = player
        self.d_state.copy_(self.h_state, non_blocking=True)


def _get() -> _Ctx:
    global _ctx
    if _ctx is None:
        _ctx = _Ctx()
    return _ctx


def legal_move_ids(state_np, player: int) -> np.ndarray:
    """Ascending move ids of one packed state for `player` (0..3): one upload, one kernel, one download."""
    with _lock:
        c = _get()
        c.h_state_np[0] = np.frombuffer(np.asarray(state_np).tobytes(), dtype=np.uint8)
        c.h_state_np[0, 240] = player                     # the player rides in the state's side-to-move byte
        # upload, kernel, download (8 KB: cheaper than a second round trip for the exact length) and the wait in ONE C call
        dev._lib.check(c.lib.bk_legal_moves_host(c.h_state.data_ptr(), c.d_state.data_ptr(), c.d_list.data_ptr(), c.h_list.data_ptr(),
                                                 MAX_MOVES, dev._stream()))
        dev._count(1)
        n = int(c.h_count_np[0])
        if n > MAX_MOVES:
            raise dev._lib.BlokusB200Error(f"{n} legal moves exceed the single-state buffer ({MAX_MOVES})")
        return c.h_ids_np[:n].copy()


def count_moves(state_np, player: int) -> int:
    with _lock:
        c = _get()
        c.upload(state_np, player)
        return int(dev.movegen_count(c.d_state, c.d_player).item())


def has_moves_mask(state_np) -> int:
    """Bit p set iff player p has a legal move."""
    with _lock:
        c = _get()
        c.upload(state_np, 0)
        return int(dev.has_moves(c.d_state).item())


def is_legal(state_np, player: int, move_id: int) -> bool:
    if not 0 <= int(move_id) < 0xFFFF:
        return False
    with _lock:
        c = _get()
        c.upload(state_np, player)
        c.h_move[0] = int(move_id)
        c.d_move.copy_(c.h_move, non_blocking=True)
        return bool(dev.is_move_legal(c.d_state, c.d_move, c.d_player).item())


def step_check(state_np, player: int, move_id: int) -> Tuple[bool, int]:
    """BlokusGame.make_move's device work in ONE launch (bk_step): is the move legal for `player`, and -- after placing
    it -- which players still have a move (bit p; 0 = game over, engine/game.py:182-214).  (False, 0) for ids out of range."""
    if not 0 <= int(move_id) < 0xFFFF:
        return False, 0
    with _lock:
        c = _get()
        c.upload_as_mover(state_np, player)
        c.h_move[0] = int(move_id)
        c.d_move.copy_(c.h_move, non_blocking=True)
        ok, alive = dev.step(c.d_state, c.d_move, None, want_alive=True)
        both = torch.stack([ok[0], alive[0]]).cpu()
        return bool(both[0]), int(both[1])


def can_place(state_np, player: int, cell_mask: int) -> bool:
    with _lock:
        c = _get()
        c.upload(state_np, player)
        c.h_mask.numpy()[0] = np.frombuffer(int(cell_mask).to_bytes(56, "little"), dtype="<i8")
        c.d_mask.copy_(c.h_mask, non_blocking=True)
        return bool(dev.can_place_mask(c.d_state, c.d_mask, c.d_player).item())


def scores(state_np) -> Tuple[List[int], List[int]]:
    """(BlokusGame.get_score per player, Board.get_score per player)."""
    with _lock:
        c = _get()
        c.upload(state_np, 0)
        gs, bs = dev.scores(c.d_state)
        both = torch.stack([gs[0], bs[0]]).cpu().tolist()
        return both[0], both[1]

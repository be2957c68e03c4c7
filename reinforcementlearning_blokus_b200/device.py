"""Batched device API: torch tensors in, torch tensors out, CUDA kernels in between.

Thin layer over the C ABI (include/blokus_b200.h).  PyTorch is used for device memory and
streams only; every function launches on torch's current stream.  States are ``uint8[N, 256]``
CUDA tensors holding packed ``bk_state`` records (see ``state.py``).
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BK_BITMAP_WORDS, BK_MAX_TURNS, STATE_BYTES, check

POLICY_UNIFORM = 0
POLICY_HEURISTIC = 1

_launches = 0  # kernels launched through this module (bench.py reports it)


def launches() -> int:
    return _launches


def _count(n: int = 1) -> None:
    global _launches
    _launches += n


def _lib_for(t: Optional[torch.Tensor] = None):
    if not torch.cuda.is_available():
        raise _lib.BlokusB200Error("no CUDA device: blokus_b200 has no CPU fallback")
    cur = torch.cuda.current_device()
    if t is not None and t.is_cuda and t.device.index != cur:
        # the C ABI launches on the caller's current device and torch's current stream belongs to it
        raise ValueError(f"tensor lives on cuda:{t.device.index} but the current device is cuda:{cur}; "
                         f"wrap the call in `with torch.cuda.device({t.device.index}):`")
    return _lib.init(cur)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw query is ~20 us cheaper per call than
    building a torch.cuda.Stream object, which matters for one-move latencies of a millisecond)."""
    if _raw_stream is not None:
        return int(_raw_stream(torch.cuda.current_device()))
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check_states(states: torch.Tensor) -> int:
    if not (states.is_cuda and states.dtype == torch.uint8 and states.dim() == 2
            and states.shape[1] == STATE_BYTES and states.is_contiguous()):
        raise ValueError("states must be a contiguous CUDA uint8 tensor of shape [N, 256]")
    if states.data_ptr() % 8:
        raise ValueError("states must be 8-byte aligned")
    return states.shape[0]


def _check_players(players: Optional[torch.Tensor], n: int) -> None:
    if players is None:
        return
    if not (players.is_cuda and players.dtype == torch.uint8 and players.numel() == n and players.is_contiguous()):
        raise ValueError("players must be a contiguous CUDA uint8 tensor of length N (values 0..3)")


def to_device(states_np: np.ndarray, device: Optional[torch.device] = None) -> torch.Tensor:
    """numpy packed states (STATE_DTYPE or uint8[N,256]) -> CUDA uint8[N,256]."""
    raw = np.ascontiguousarray(states_np).view(np.uint8).reshape(-1, STATE_BYTES)
    return torch.from_numpy(raw.copy()).to(device or torch.device("cuda", torch.cuda.current_device()))


# ---- K1 --------------------------------------------------------------------------------------
def movegen_count(states: torch.Tensor, players: Optional[torch.Tensor] = None) -> torch.Tensor:
    """len(get_legal_moves(board, player)) per state (engine/move_generator.py:130)."""
    n = _check_states(states)
    _check_players(players, n)
    lib = _lib_for(states)
    counts = torch.empty(n, dtype=torch.int32, device=states.device)
    check(lib.bk_movegen_count(_ptr(states), _ptr(players), n, _ptr(counts), _stream()))
    _count()
    return counts


ONE_PASS_CAP = 1280          # row capacity of the one-pass CSR path (longest list seen in 10^7 states: 1 052)
ONE_PASS_MAX_BYTES = 6 << 30  # temporary rows beyond this size: use the two-pass path


def movegen(states: torch.Tensor, players: Optional[torch.Tensor] = None,
            one_pass: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Legal moves of every state as CSR: (counts int32[N], offsets int32[N+1], move_ids uint16[total]),
    ids ascending = the (piece, orientation, row, col) order of the reference's naive generator.

    one_pass (default: whenever the temporary rows fit ONE_PASS_MAX_BYTES): ONE generation per state writes ids into
    fixed rows of ONE_PASS_CAP slots, a scan of the counts gives the offsets and a copy kernel packs the rows; a list
    longer than the cap (never seen) falls back to the two-pass path (count, scan, second generation that emits)."""
    n = _check_states(states)
    _check_players(players, n)
    lib = _lib_for(states)
    st = _stream()
    offsets = torch.empty(n + 1, dtype=torch.int32, device=states.device)
    if one_pass is None:
        one_pass = n * ONE_PASS_CAP * 2 <= ONE_PASS_MAX_BYTES
    if one_pass and n > 0:
        counts, rows = movegen_fixed(states, ONE_PASS_CAP, players)
        check(lib.bk_exclusive_scan_u32(_ptr(counts), n, _ptr(offsets), st))
        _count(3)
        total, longest = (int(v) for v in torch.stack([counts.sum(dtype=torch.int64), counts.max().to(torch.int64)]).tolist())
        _check_total(total)
        ids = torch.empty(max(total, 1), dtype=torch.uint16, device=states.device)
        if longest <= ONE_PASS_CAP:
            check(lib.bk_compact_rows(_ptr(rows), ONE_PASS_CAP, _ptr(offsets), n, _ptr(ids), st))
        else:
            check(lib.bk_movegen_emit(_ptr(states), _ptr(players), n, _ptr(offsets), _ptr(ids), st))
        _count()
        return counts, offsets, ids[:total]
    counts = torch.empty(n, dtype=torch.int32, device=states.device)
    check(lib.bk_movegen_count(_ptr(states), _ptr(players), n, _ptr(counts), st))
    check(lib.bk_exclusive_scan_u32(_ptr(counts), n, _ptr(offsets), st))
    _count(4)
    total = int(counts.sum(dtype=torch.int64).item())     # not offsets[n]: the 32-bit scan wraps on huge batches
    _check_total(total)
    ids = torch.empty(max(total, 1), dtype=torch.uint16, device=states.device)
    check(lib.bk_movegen_emit(_ptr(states), _ptr(players), n, _ptr(offsets), _ptr(ids), st))
    _count()
    return counts, offsets, ids[:total]


def _check_total(total: int) -> None:
    """CSR offsets are 32-bit: refuse a batch whose move ids do not fit instead of writing out of bounds."""
    if total >= 2**31:
        raise ValueError(f"the batch holds {total} legal moves; CSR offsets are int32 -- split it into batches below 2^31 ids "
                         "(about 10 M mid-game states)")


def exclusive_scan(counts: torch.Tensor) -> torch.Tensor:
    """int32[n] -> int32[n+1] exclusive prefix sums (last entry = total); three short kernels, no allocation."""
    n = counts.numel()
    lib = _lib_for(counts)
    offsets = torch.empty(n + 1, dtype=torch.int32, device=counts.device)
    check(lib.bk_exclusive_scan_u32(_ptr(counts), n, _ptr(offsets), _stream()))
    _count(3)
    return offsets


def movegen_fixed(states: torch.Tensor, cap: int, players: Optional[torch.Tensor] = None,
                  out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One-kernel generation into fixed slots: (counts int32[N], ids uint16[N, cap]); a list longer than
    cap is cut (counts still holds the full length)."""
    n = _check_states(states)
    _check_players(players, n)
    lib = _lib_for(states)
    if out is None:
        counts = torch.empty(n, dtype=torch.int32, device=states.device)
        ids = torch.empty((n, cap), dtype=torch.uint16, device=states.device)
    else:
        counts, ids = out
    check(lib.bk_movegen_fixed(_ptr(states), _ptr(players), n, int(cap), _ptr(counts), _ptr(ids), _stream()))
    _count()
    return counts, ids


def can_place_mask(states: torch.Tensor, cell_masks: torch.Tensor, players: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Board.can_place_piece for arbitrary cell sets given as 400-bit masks int64[N,7] (engine/board.py:136)."""
    n = _check_states(states)
    _check_players(players, n)
    if not (cell_masks.is_cuda and cell_masks.dtype in (torch.int64, torch.uint64) and tuple(cell_masks.shape) == (n, 7)
            and cell_masks.is_contiguous()):
        raise ValueError("cell_masks must be a contiguous CUDA 64-bit tensor of shape [N, 7]")
    lib = _lib_for(states)
    out = torch.empty(n, dtype=torch.uint8, device=states.device)
    check(lib.bk_can_place_mask(_ptr(states), _ptr(players), _ptr(cell_masks), n, _ptr(out), _stream()))
    _count()
    return out


def movegen_bitmap(states: torch.Tensor, players: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int32[N, 91*20]: word o*20+row has bit col set iff move (o,row,col) is legal."""
    n = _check_states(states)
    _check_players(players, n)
    lib = _lib_for(states)
    bitmap = torch.empty((n, BK_BITMAP_WORDS), dtype=torch.int32, device=states.device)
    check(lib.bk_movegen_bitmap(_ptr(states), _ptr(players), n, _ptr(bitmap), _stream()))
    _count()
    return bitmap


def piece_counts(states: torch.Tensor, players: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(P int16[N, 21], anchor_max int16[N]): legal placements per piece and the most moves sharing one anchor cell
    (inputs of engine/mobility_metrics.py:38-131)."""
    n = _check_states(states)
    _check_players(players, n)
    lib = _lib_for(states)
    pc = torch.empty((n, 21), dtype=torch.int16, device=states.device)
    am = torch.empty(n, dtype=torch.int16, device=states.device)
    check(lib.bk_movegen_piece_counts(_ptr(states), _ptr(players), n, _ptr(pc), _ptr(am), _stream()))
    _count()
    return pc, am


NUM_SNAPSHOT_FEATURES = 34


def snapshot_features(states: torch.Tensor, frontiers: Optional[torch.Tensor] = None,
                      turn_index: Optional[torch.Tensor] = None, max_turns: int = 2500) -> torch.Tensor:
    """float64[N, 4, 34]: the SNAPSHOT_FEATURE_COLUMNS rows of all four players of every state
    (analytics/winprob/features.py:103-193).  frontiers: int64[N, 4, 7] bitboards of the reference's incremental
    frontier sets, or None for the sets of a freshly restored board; turn_index: int32[N] or None (= move_count)."""
    n = _check_states(states)
    if frontiers is not None and not (frontiers.is_cuda and frontiers.dtype == torch.int64 and frontiers.is_contiguous()
                                      and tuple(frontiers.shape) == (n, 4, 7)):
        raise ValueError("frontiers must be a contiguous CUDA int64 tensor of shape [N, 4, 7]")
    if turn_index is not None and not (turn_index.is_cuda and turn_index.dtype == torch.int32 and turn_index.is_contiguous()
                                       and turn_index.numel() == n):
        raise ValueError("turn_index must be a contiguous CUDA int32 tensor of length N")
    lib = _lib_for(states)
    out = torch.empty((n, 4, NUM_SNAPSHOT_FEATURES), dtype=torch.float64, device=states.device)
    check(lib.bk_snapshot_features(_ptr(states), _ptr(frontiers), _ptr(turn_index), n, int(max_turns), _ptr(out), _stream()))
    _count()
    return out


def replay_snapshots(traces: torch.Tensor, checkpoints, max_turns: int = 2500, start: Optional[torch.Tensor] = None,
                     want_states: bool = False) -> dict:
    """Replay recorded games (uint16[N, T] move ids, 0xFFFF = pass) and capture the feature rows at the checkpoint
    plies, where analytics/tournament/arena_runner.py:625-650,698 captures them.  Returns features float64[N, K, 4, 34]
    (zero where not captured), turn_index int32[N, K] (-1 = game never reached that ply) and, with want_states, the
    captured boards (states uint8[N, K, 256], frontiers int64[N, K, 4, 7])."""
    if not (traces.is_cuda and traces.dtype == torch.uint16 and traces.dim() == 2 and traces.is_contiguous()):
        raise ValueError("traces must be a contiguous CUDA uint16 tensor of shape [N, T]")
    n, n_turns = traces.shape
    ck = np.ascontiguousarray(sorted({int(c) for c in checkpoints}), dtype=np.int32)
    if not 1 <= len(ck) <= 32 or ck.min() < 0:
        raise ValueError("1..32 distinct non-negative checkpoint plies required")
    if start is not None and _check_states(start) != n:
        raise ValueError("start must hold one state per trace")
    lib = _lib_for(traces)
    k = len(ck)
    feats = torch.zeros((n, k, 4, NUM_SNAPSHOT_FEATURES), dtype=torch.float64, device=traces.device)
    ti = torch.empty((n, k), dtype=torch.int32, device=traces.device)
    st = torch.zeros((n, k, STATE_BYTES), dtype=torch.uint8, device=traces.device) if want_states else None
    fr = torch.zeros((n, k, 4, 7), dtype=torch.int64, device=traces.device) if want_states else None
    check(lib.bk_replay_snapshots(_ptr(start), _ptr(traces), n, n_turns, ck.ctypes.data, k, int(max_turns), _ptr(feats),
                                  _ptr(ti), _ptr(st), _ptr(fr), _stream()))
    _count()
    res = {"features": feats, "turn_index": ti, "checkpoints": ck}
    if want_states:
        res["states"], res["frontiers"] = st, fr
    return res


def has_moves(states: torch.Tensor) -> torch.Tensor:
    """uint8[N]: bit p set iff player p has a legal move (engine/move_generator.py:961)."""
    n = _check_states(states)
    lib = _lib_for(states)
    out = torch.empty(n, dtype=torch.uint8, device=states.device)
    check(lib.bk_has_moves(_ptr(states), n, _ptr(out), _stream()))
    _count()
    return out


def is_move_legal(states: torch.Tensor, move_ids: torch.Tensor, players: Optional[torch.Tensor] = None) -> torch.Tensor:
    n = _check_states(states)
    _check_players(players, n)
    if not (move_ids.is_cuda and move_ids.dtype == torch.uint16 and move_ids.numel() == n):
        raise ValueError("move_ids must be a CUDA uint16 tensor of length N")
    lib = _lib_for(states)
    out = torch.empty(n, dtype=torch.uint8, device=states.device)
    check(lib.bk_is_move_legal(_ptr(states), _ptr(players), _ptr(move_ids), n, _ptr(out), _stream()))
    _count()
    return out


def step(states: torch.Tensor, move_ids: torch.Tensor, players: Optional[torch.Tensor] = None,
         want_alive: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """In-place make_move / pass per state (engine/game.py:57, engine/board.py:557).  Returns (ok, alive)."""
    n = _check_states(states)
    _check_players(players, n)
    if not (move_ids.is_cuda and move_ids.dtype == torch.uint16 and move_ids.numel() == n):
        raise ValueError("move_ids must be a CUDA uint16 tensor of length N")
    lib = _lib_for(states)
    ok = torch.empty(n, dtype=torch.uint8, device=states.device)
    alive = torch.empty(n, dtype=torch.uint8, device=states.device) if want_alive else None
    check(lib.bk_step(_ptr(states), _ptr(players), _ptr(move_ids), n, _ptr(ok), _ptr(alive), _stream()))
    _count()
    return ok, alive


def scores(states: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(BlokusGame.get_score, Board.get_score) as int32[N,4] each (engine/game.py:284, engine/board.py:562)."""
    n = _check_states(states)
    lib = _lib_for(states)
    gs = torch.empty((n, 4), dtype=torch.int32, device=states.device)
    bs = torch.empty((n, 4), dtype=torch.int32, device=states.device)
    check(lib.bk_scores(_ptr(states), n, _ptr(gs), _ptr(bs), _stream()))
    _count()
    return gs, bs


# ---- K2 --------------------------------------------------------------------------------------
def playout(n: int, seed: int, game_offset: int = 0, start_states: Optional[torch.Tensor] = None,
            policy: int = POLICY_UNIFORM, seat_policies: Optional[torch.Tensor] = None, max_turns: int = 0, first_turn: int = 0, want_trace: bool = False,
            want_choice: bool = False, want_final: bool = False, want_board_scores: bool = True,
            out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """Play n games to the end on the device.  Game i has global index game_offset+i (its Philox stream).
    Turns first_turn .. max_turns-1 are played (0 = to the end); segments chained through final_states give
    the same game as one call.  `seat_policies` (int32[N], policy of player p in byte p) overrides `policy` per
    seat and game (arena match-ups).

    `out` may carry preallocated tensors from a previous call (keys as returned) to avoid reallocation.
    """
    if start_states is not None:
        if _check_states(start_states) != n:
            raise ValueError("start_states must have N rows")
    lib = _lib_for(start_states)
    dev = start_states.device if start_states is not None else torch.device("cuda", torch.cuda.current_device())
    res = out if out is not None else {}

    def buf(key, shape, dtype, want=True):
        if not want:
            return None
        t = res.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev:
            t = torch.empty(shape, dtype=dtype, device=dev)
            res[key] = t
        return t

    gs = buf("game_scores", (n, 4), torch.int32)
    bs = buf("board_scores", (n, 4), torch.int32, want_board_scores)
    tr = buf("trace", (n, BK_MAX_TURNS), torch.uint16, want_trace)
    ch = buf("choice", (n, BK_MAX_TURNS), torch.uint16, want_choice)
    fin = buf("final_states", (n, STATE_BYTES), torch.uint8, want_final)
    if seat_policies is not None:
        if not (seat_policies.is_cuda and seat_policies.dtype == torch.int32 and seat_policies.numel() == n):
            raise ValueError("seat_policies must be a CUDA int32 tensor of length N (policy of player p in byte p)")
        check(lib.bk_playout_seats(_ptr(start_states), n, int(seed) & (2**64 - 1), int(game_offset), _ptr(seat_policies),
                                   int(first_turn), int(max_turns), _ptr(gs), _ptr(bs), _ptr(tr), _ptr(ch), _ptr(fin),
                                   _stream()))
    else:
        check(lib.bk_playout(_ptr(start_states), n, int(seed) & (2**64 - 1), int(game_offset), int(policy),
                             int(first_turn), int(max_turns), _ptr(gs), _ptr(bs), _ptr(tr), _ptr(ch), _ptr(fin), _stream()))
    _count()
    return res


def heuristic_scores(states: torch.Tensor, offsets: torch.Tensor, move_ids: torch.Tensor,
                     players: Optional[torch.Tensor] = None, weights: Optional[Sequence[float]] = None) -> torch.Tensor:
    """float64 HeuristicAgent._evaluate_move score of every listed move (agents/heuristic_agent.py:68-103).  weights =
    (piece_size, corner_creation, edge_avoidance, center_preference) of set_weights; None = the defaults 1, 2, -1.5, 0.5."""
    n = _check_states(states)
    _check_players(players, n)
    lib = _lib_for(states)
    out = torch.empty(max(move_ids.numel(), 1), dtype=torch.float64, device=states.device)
    if weights is not None:
        w = (C.c_double * 4)(*[float(v) for v in weights])
        check(lib.bk_heuristic_scores_weighted(_ptr(states), _ptr(players), _ptr(offsets), _ptr(move_ids), n, w, _ptr(out), _stream()))
        _count()
        return out[:move_ids.numel()]
    check(lib.bk_heuristic_scores(_ptr(states), _ptr(players), _ptr(offsets), _ptr(move_ids), n, _ptr(out), _stream()))
    _count()
    return out[:move_ids.numel()]


class PlayoutPipeline:
    """Host-to-host playouts with the copies overlapped with compute: start states arrive in pinned host memory,
    final scores are wanted in pinned host memory.  Two slots, each with its own stream; while slot A's kernel runs,
    slot B's states are uploading and the previous scores are downloading.

        pipe = PlayoutPipeline(n)
        for k, host_states in enumerate(batches):          # pinned uint8[n, 256]
            pipe.submit(host_states, seed, game_offset=k * n)
            scores = pipe.collect()                          # pinned int32[n, 4] of the OLDEST batch in flight, or None
        scores = pipe.drain()                                # remaining batches, oldest first
    """

    def __init__(self, n: int, depth: int = 2, policy: int = POLICY_UNIFORM):
        self.n, self.depth, self.policy = int(n), int(depth), int(policy)
        d = torch.device("cuda", torch.cuda.current_device())
        self.slots = []
        for _ in range(self.depth):
            self.slots.append({
                "stream": torch.cuda.Stream(device=d),
                "dstates": torch.empty((n, STATE_BYTES), dtype=torch.uint8, device=d),
                "out": {},
                "hscores": torch.empty((n, 4), dtype=torch.int32).pin_memory(),
                "done": torch.cuda.Event(),
                "busy": False,
            })
        self._head = 0      # next slot to fill
        self._tail = 0      # oldest slot in flight
        self._inflight = 0

    def submit(self, host_states: torch.Tensor, seed: int, game_offset: int = 0) -> None:
        if self._inflight == self.depth:
            raise RuntimeError("pipeline full: collect() a result first")
        slot = self.slots[self._head]
        with torch.cuda.stream(slot["stream"]):
            slot["dstates"].copy_(host_states, non_blocking=True)
            r = playout(self.n, seed=seed, game_offset=game_offset, start_states=slot["dstates"], policy=self.policy,
                        out=slot["out"], want_board_scores=False)
            slot["hscores"].copy_(r["game_scores"], non_blocking=True)
            slot["done"].record(slot["stream"])
        slot["busy"] = True
        self._head = (self._head + 1) % self.depth
        self._inflight += 1

    def collect(self, wait_if_full_only: bool = True):
        """Scores of the oldest batch in flight.  With wait_if_full_only the call returns None while the pipeline
        still has a free slot (so the caller keeps it full); otherwise it waits for the oldest batch."""
        if self._inflight == 0 or (wait_if_full_only and self._inflight < self.depth):
            return None
        slot = self.slots[self._tail]
        slot["done"].synchronize()
        slot["busy"] = False
        self._tail = (self._tail + 1) % self.depth
        self._inflight -= 1
        return slot["hscores"]

    def drain(self):
        out = []
        while self._inflight:
            out.append(self.collect(wait_if_full_only=False).clone())
        return out


# ---- K4 --------------------------------------------------------------------------------------
def zobrist_hash(states: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """uint64 keys (returned as int64 bit patterns) of ZobristHash.hash_board (mcts/zobrist.py:70)."""
    n = _check_states(states)
    if not (table.is_cuda and table.dtype in (torch.int64, torch.uint64) and table.numel() == _lib.BK_ZOBRIST_ENTRIES):
        raise ValueError("table must be a CUDA 64-bit tensor with 2088 entries")
    lib = _lib_for(states)
    keys = torch.empty(n, dtype=torch.int64, device=states.device)
    check(lib.bk_zobrist_hash(_ptr(states), n, _ptr(table), _ptr(keys), _stream()))
    _count()
    return keys


class DeviceTranspositionTable:
    """Open-addressing device table key -> reward (mcts/zobrist.py:155-220 TranspositionTable)."""

    def __init__(self, log2_slots: int = 20):
        import ctypes as C

        self._lib = _lib_for(None)
        h = C.c_void_p()
        check(self._lib.bk_tt_create(C.byref(h), int(log2_slots)))
        self._h = h
        self.log2_slots = int(log2_slots)

    @property
    def handle(self):
        return self._h

    def clear(self) -> None:
        check(self._lib.bk_tt_clear(self._h, _stream()))

    def put(self, keys: torch.Tensor, rewards: torch.Tensor) -> None:
        n = keys.numel()
        rewards = rewards.to(torch.float64).contiguous()
        check(self._lib.bk_tt_put(self._h, _ptr(keys), _ptr(rewards), n, _stream()))
        _count()

    def get(self, keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        n = keys.numel()
        rewards = torch.zeros(n, dtype=torch.float64, device=keys.device)
        found = torch.empty(n, dtype=torch.uint8, device=keys.device)
        check(self._lib.bk_tt_get(self._h, _ptr(keys), n, _ptr(rewards), _ptr(found), _stream()))
        _count()
        return rewards, found

    def stats(self) -> Dict[str, int]:
        import ctypes as C

        out = (C.c_uint64 * 4)()
        check(self._lib.bk_tt_stats(self._h, out, _stream()))
        return {"max_size": int(out[0]), "size": int(out[1]), "hits": int(out[2]), "probes": int(out[3])}

    def evict_oldest(self, n: int) -> None:
        """Remove the n entries inserted first (TranspositionTable.put's eviction, mcts/zobrist.py:196-200)."""
        check(self._lib.bk_tt_evict_oldest(self._h, int(n), _stream()))
        _count(2)

    def close(self) -> None:
        if self._h is not None:
            self._lib.bk_tt_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

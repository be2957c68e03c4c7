"""ctypes front-end of the CPU oracle (``blokus_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of ``blokus_oracle.c``.  Imported by ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``; never by the product
package.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")

PASS_ID = 0xFFFF
MAX_TURNS = 96
STATE_BYTES = 256

STATE_DTYPE = np.dtype(
    [
        ("bits", "<u8", (4, 7)),
        ("used", "<u4", (4,)),
        ("to_move", "u1"),
        ("first_mask", "u1"),
        ("move_count", "<u2"),
        ("reserved", "u1", (12,)),
    ]
)
assert STATE_DTYPE.itemsize == STATE_BYTES


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "blokus_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, u32, u64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64
        L.orc_init.restype = None
        L.orc_num_orientations.restype = i32
        L.orc_orientation_info.argtypes = [i32, vp]
        L.orc_board_sizeof.restype = i32
        L.orc_board_reset.argtypes = [vp]
        L.orc_can_place.argtypes = [vp, i32, vp, vp, i32]
        L.orc_pass.argtypes = [vp]
        L.orc_place.argtypes = [vp, i32, i32, i32, i32, i32]
        L.orc_legal_moves_naive.argtypes = [vp, i32, vp, i32]
        L.orc_legal_moves_frontier.argtypes = [vp, i32, vp, i32]
        L.orc_has_legal_moves.argtypes = [vp, i32]
        L.orc_is_move_legal.argtypes = [vp, i32, i32]
        L.orc_check_game_over.argtypes = [vp]
        L.orc_make_move.argtypes = [vp, i32, i32]
        L.orc_board_score.argtypes = [vp, i32]
        L.orc_game_score.argtypes = [vp, i32]
        L.orc_pack.argtypes = [vp, vp]
        L.orc_unpack.argtypes = [vp, vp]
        L.orc_get_grid.argtypes = [vp, vp]
        L.orc_get_current_player.argtypes = [vp]
        L.orc_get_game_over.argtypes = [vp]
        L.orc_get_move_count.argtypes = [vp]
        L.orc_get_used.argtypes = [vp, i32]
        L.orc_get_used.restype = u32
        L.orc_get_frontier.argtypes = [vp, i32, vp]
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        L.orc_heuristic_score.argtypes = [vp, i32, i32]
        L.orc_heuristic_score.restype = C.c_double
        L.orc_playout.argtypes = [vp, u64, u64, i32, i32, i32, i32, vp, vp, vp, vp, vp]
        L.orc_zobrist_table.argtypes = [u32, vp]
        L.orc_zobrist_hash.argtypes = [vp, vp]
        L.orc_zobrist_hash.restype = u64
        L.orc_tt_create.argtypes = [i32]
        L.orc_tt_create.restype = vp
        L.orc_tt_destroy.argtypes = [vp]
        L.orc_tt_size.argtypes = [vp]
        L.orc_tt_size.restype = u64
        for name in ("orc_tt_hits", "orc_tt_probes"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = u64
        L.orc_tt_get.argtypes = [vp, u64, vp]
        L.orc_tt_put.argtypes = [vp, u64, C.c_double]
        L.orc_tt_put.restype = None
        L.orc_tt_evict_oldest.argtypes = [vp, u64]
        L.orc_tt_evict_oldest.restype = None
        L.orc_set_tt_namespace.argtypes = [i32]
        L.orc_mcts_search.argtypes = [vp, i32, i32, i32, i32, i32, i32, C.c_double, u64, u64, vp, vp, i32, vp, vp, vp]
        L.orc_mcts_search_learned.argtypes = [vp, vp, i32, i32, i32, C.c_double, u64, u64, vp, vp, i32, vp, i32, C.c_double, vp, vp, vp]
        L.orc_set_model_kind.argtypes = [i32]
        L.orc_win_probability_packed.argtypes = [vp, vp, i32, vp, i32, i32]
        L.orc_win_probability_packed.restype = C.c_double
        L.orc_mcts_search_shaped.argtypes = [vp, vp, i32, i32, i32, C.c_double, u64, u64, vp, vp, i32, vp, i32, C.c_double, C.c_double,
                                             C.c_double, i32, vp, vp, vp]
        L.orc_mcts_search_batched.argtypes = [vp, i32, i32, i32, i32, i32, i32, C.c_double, u64, u64, vp, vp, i32, i32, vp, vp, vp]
        L.orc_mcts_wide.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, C.c_double, u64, u64, i32, vp, vp, vp, vp]
        L.orc_check_moves_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp]
        L.orc_check_playouts.argtypes = [u64, u64, i32, i32, i32, vp, vp, vp]
        L.orc_num_snapshot_features.restype = i32
        L.orc_snapshot_features.argtypes = [vp, i32, i32, vp]
        L.orc_snapshot_features_packed.argtypes = [vp, vp, i32, i32, vp]
        L.orc_replay_snapshots.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, vp, vp]
        L.orc_init()
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleBoard:
    """A reference-rules board living in C memory (restates engine/board.py:43-78)."""

    def __init__(self):
        L = lib()
        self._buf = C.create_string_buffer(L.orc_board_sizeof())
        L.orc_board_reset(self._buf)

    # -- construction -----------------------------------------------------------------------
    @classmethod
    def from_packed(cls, state: np.ndarray) -> "OracleBoard":
        b = cls()
        s = np.ascontiguousarray(state, dtype=STATE_DTYPE).reshape(())
        lib().orc_unpack(_ptr(np.atleast_1d(s)), b._buf)
        return b

    def copy(self) -> "OracleBoard":
        b = OracleBoard.__new__(OracleBoard)
        b._buf = C.create_string_buffer(self._buf.raw, len(self._buf))
        return b

    def pack(self) -> np.ndarray:
        out = np.zeros(1, dtype=STATE_DTYPE)
        lib().orc_pack(self._buf, _ptr(out))
        return out[0]

    # -- state getters ----------------------------------------------------------------------
    @property
    def grid(self) -> np.ndarray:
        g = np.zeros((20, 20), dtype=np.int8)
        lib().orc_get_grid(self._buf, _ptr(g))
        return g

    @property
    def current_player(self) -> int:
        return lib().orc_get_current_player(self._buf)

    @property
    def game_over(self) -> bool:
        return bool(lib().orc_get_game_over(self._buf))

    @property
    def move_count(self) -> int:
        return lib().orc_get_move_count(self._buf)

    def used(self, player: int) -> int:
        return lib().orc_get_used(self._buf, player)

    def frontier(self, player: int) -> List[Tuple[int, int]]:
        f = np.zeros(400, dtype=np.uint8)
        lib().orc_get_frontier(self._buf, player, _ptr(f))
        return [(int(i) // 20, int(i) % 20) for i in np.nonzero(f)[0]]

    # -- rules ------------------------------------------------------------------------------
    def legal_moves(self, player: int, naive: bool = False) -> np.ndarray:
        out = np.zeros(4096, dtype=np.uint16)
        fn = lib().orc_legal_moves_naive if naive else lib().orc_legal_moves_frontier
        n = fn(self._buf, player, _ptr(out), out.size)
        assert n <= out.size
        return out[:n].copy()

    def has_legal_moves(self, player: int) -> bool:
        return bool(lib().orc_has_legal_moves(self._buf, player))

    def is_move_legal(self, player: int, move_id: int) -> bool:
        return bool(lib().orc_is_move_legal(self._buf, player, int(move_id)))

    def can_place(self, player: int, cells: Sequence[Tuple[int, int]]) -> bool:
        r = np.array([c[0] for c in cells], dtype=np.int32)
        c = np.array([c[1] for c in cells], dtype=np.int32)
        return bool(lib().orc_can_place(self._buf, player, _ptr(r), _ptr(c), len(cells)))

    def place(self, player: int, move_id: int, validate: bool = True) -> bool:
        g, cell = divmod(int(move_id), 400)
        return bool(lib().orc_place(self._buf, player, g, cell // 20, cell % 20, int(validate)))

    def make_move(self, player: int, move_id: int) -> bool:
        return bool(lib().orc_make_move(self._buf, player, int(move_id)))

    def pass_turn(self) -> None:
        lib().orc_pass(self._buf)

    def check_game_over(self) -> bool:
        lib().orc_check_game_over(self._buf)
        return self.game_over

    def board_score(self, player: int) -> int:
        return lib().orc_board_score(self._buf, player)

    def game_score(self, player: int) -> int:
        return lib().orc_game_score(self._buf, player)

    def heuristic_score(self, player: int, move_id: int) -> float:
        return lib().orc_heuristic_score(self._buf, player, int(move_id))

    def zobrist_hash(self, table: np.ndarray) -> int:
        t = np.ascontiguousarray(table, dtype=np.uint64)
        return int(lib().orc_zobrist_hash(_ptr(t), self._buf))


def orientation_table() -> List[dict]:
    L = lib()
    out = []
    for g in range(L.orc_num_orientations()):
        info = np.zeros(15, dtype=np.int32)
        L.orc_orientation_info(g, _ptr(info))
        n = int(info[2])
        out.append(
            {
                "piece_id": int(info[0]),
                "index": int(info[1]),
                "h": int(info[3]),
                "w": int(info[4]),
                "cells": [(int(info[5 + 2 * i]), int(info[6 + 2 * i])) for i in range(n)],
            }
        )
    return out


def philox4x32_10(ctr: Sequence[int], key: Sequence[int]) -> Tuple[int, int, int, int]:
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_ptr(c), _ptr(k), _ptr(o))
    return tuple(int(x) for x in o)


def zobrist_table(seed: int) -> np.ndarray:
    t = np.zeros(2088, dtype=np.uint64)
    lib().orc_zobrist_table(int(seed) & 0xFFFFFFFF, _ptr(t))
    return t


def playout(
    seed: int,
    game_index: int,
    start: Optional[np.ndarray] = None,
    max_turns: int = 0,
    fast: bool = True,
    first_turn: int = 0,
    policy: int = 0,
    want_final: bool = False,
):
    """One Philox-driven random game; returns dict(game_scores, board_scores, trace, choice, turns)."""
    gs = np.zeros(4, dtype=np.int32)
    bs = np.zeros(4, dtype=np.int32)
    tr = np.full(MAX_TURNS, PASS_ID, dtype=np.uint16)
    ch = np.full(MAX_TURNS, PASS_ID, dtype=np.uint16)
    fin = np.zeros(1, dtype=STATE_DTYPE) if want_final else None
    st = None
    if start is not None:
        st = np.ascontiguousarray(np.atleast_1d(start), dtype=STATE_DTYPE)
    turns = lib().orc_playout(_ptr(st), int(seed), int(game_index), int(policy), int(first_turn), int(max_turns), int(fast),
                              _ptr(gs), _ptr(bs), _ptr(tr), _ptr(ch), _ptr(fin))
    res = {"game_scores": gs, "board_scores": bs, "trace": tr, "choice": ch, "turns": int(turns)}
    if want_final:
        res["final"] = fin[0]
    return res


NUM_SNAPSHOT_FEATURES = 34


def snapshot_features(state: np.ndarray, turn_index: Optional[int] = None, max_turns: int = 2500,
                      frontiers: Optional[np.ndarray] = None) -> np.ndarray:
    """float64[4, 34] feature rows of one packed state (analytics/winprob/features.py).  frontiers = uint8[4, 400]
    incremental frontier sets, or None for the sets rebuilt from the cells; turn_index None = move_count."""
    st = np.ascontiguousarray(np.atleast_1d(state), dtype=STATE_DTYPE)
    out = np.zeros((4, NUM_SNAPSHOT_FEATURES), dtype=np.float64)
    fr = None if frontiers is None else np.ascontiguousarray(frontiers, dtype=np.uint8)
    ti = int(st[0]["move_count"]) if turn_index is None else int(turn_index)
    lib().orc_snapshot_features_packed(_ptr(st), _ptr(fr), ti, int(max_turns), _ptr(out))
    return out


def replay_snapshots(trace: np.ndarray, checkpoints: Sequence[int], max_turns: int = 2500, start: Optional[np.ndarray] = None):
    """Replay one recorded game and capture the feature rows at the checkpoint plies (arena_runner.py:625-650,698).
    Returns dict(features float64[k,4,34], turn_index int32[k] (-1 = never reached), states, frontiers uint8[k,4,400])."""
    tr = np.ascontiguousarray(trace, dtype=np.uint16)
    ck = np.ascontiguousarray(checkpoints, dtype=np.int32)
    k = len(ck)
    feats = np.zeros((k, 4, NUM_SNAPSHOT_FEATURES), dtype=np.float64)
    ti = np.zeros(k, dtype=np.int32)
    states = np.zeros(k, dtype=STATE_DTYPE)
    fr = np.zeros((k, 4, 400), dtype=np.uint8)
    st = None if start is None else np.ascontiguousarray(np.atleast_1d(start), dtype=STATE_DTYPE)
    lib().orc_replay_snapshots(_ptr(st), _ptr(tr), len(tr), _ptr(ck), k, int(max_turns), _ptr(feats), _ptr(ti), _ptr(states), _ptr(fr))
    return {"features": feats, "turn_index": ti, "states": states, "frontiers": fr}


class OracleTT:
    """Host transposition table for the oracle UCT (mcts/zobrist.py:155-220)."""

    def __init__(self, log2_slots: int = 20):
        self.h = lib().orc_tt_create(int(log2_slots))

    def size(self) -> int:
        return int(lib().orc_tt_size(self.h))

    def hits(self) -> int:
        """hit_count of the searches that used the table (mcts_agent.py:423)."""
        return int(lib().orc_tt_hits(self.h))

    def probes(self) -> int:
        return int(lib().orc_tt_probes(self.h))

    def get(self, key: int):
        v = C.c_double()
        return float(v.value) if lib().orc_tt_get(self.h, int(key), C.byref(v)) else None

    def put(self, key: int, reward: float) -> None:
        lib().orc_tt_put(self.h, int(key), float(reward))

    def evict_oldest(self, n: int) -> None:
        """TranspositionTable.put's eviction (mcts/zobrist.py:196-200): the n entries inserted first go."""
        lib().orc_tt_evict_oldest(self.h, int(n))

    def close(self):
        if self.h:
            lib().orc_tt_destroy(self.h)
            self.h = None


def mcts_search(root: np.ndarray, iterations: int, seed: int, tree_id: int = 0, max_rollout_moves: int = 50,
                max_depth: int = 0, ucb_variant: int = 0, reward_mode: int = 0, rollout_policy: int = 0, exploration: float = 1.414, tt: Optional[OracleTT] = None,
                zobrist: Optional[np.ndarray] = None, max_root_moves: int = 2048, leaf_batch: int = 1, tt_namespace: int = 0):
    """Sequential UCT with canonical untried order and Philox rollouts; returns dict(n_root, visits, reward, moves).
    tt_namespace: the tree's namespace in the transposition table (the kernel's tt_tree_base + tree index; 0 = plain keys)."""
    lib().orc_set_tt_namespace(int(tt_namespace))       # thread-local: applies to the call below
    st = np.ascontiguousarray(np.atleast_1d(root), dtype=STATE_DTYPE)
    vis = np.zeros(max_root_moves, dtype=np.int32)
    rew = np.zeros(max_root_moves, dtype=np.int64)
    mov = np.zeros(max_root_moves, dtype=np.uint16)
    zt = None if zobrist is None else np.ascontiguousarray(zobrist, dtype=np.uint64)
    n = lib().orc_mcts_search_batched(_ptr(st), int(iterations), int(max_rollout_moves), int(max_depth), int(ucb_variant),
                                      int(reward_mode), int(rollout_policy), float(exploration), int(seed), int(tree_id),
                                      tt.h if tt else None, _ptr(zt), int(max_root_moves), int(leaf_batch), _ptr(vis), _ptr(rew), _ptr(mov))
    return {"n_root": int(n), "visits": vis, "reward": rew, "moves": mov}


def mcts_search_learned(root: np.ndarray, iterations: int, seed: int, model: np.ndarray, tree_id: int = 0, max_turns: int = 2500,
                        root_frontier: Optional[np.ndarray] = None, max_depth: int = 0, ucb_variant: int = 0,
                        exploration: float = 1.414, tt: Optional[OracleTT] = None, zobrist: Optional[np.ndarray] = None,
                        max_root_moves: int = 2048, bias_weight: float = 0.0, model_kind: int = 0, tt_namespace: int = 0):
    """UCT with learned leaf evaluation (mcts_agent.py:427-451): model = float64[35] (34 weights + bias of a folded
    pairwise logistic model), root_frontier = uint8[4, 400] incremental frontier sets or None.  reward sums are doubles."""
    st = np.ascontiguousarray(np.atleast_1d(root), dtype=STATE_DTYPE)
    md = np.ascontiguousarray(model, dtype=np.float64)
    assert model_kind == 1 or md.size == 35
    lib().orc_set_model_kind(int(model_kind))           # thread-local: applies to the call below
    lib().orc_set_tt_namespace(int(tt_namespace))
    fr = None if root_frontier is None else np.ascontiguousarray(root_frontier, dtype=np.uint8)
    vis = np.zeros(max_root_moves, dtype=np.int32)
    rew = np.zeros(max_root_moves, dtype=np.float64)
    mov = np.zeros(max_root_moves, dtype=np.uint16)
    zt = None if zobrist is None else np.ascontiguousarray(zobrist, dtype=np.uint64)
    n = lib().orc_mcts_search_learned(_ptr(st), _ptr(fr), int(iterations), int(max_depth), int(ucb_variant), float(exploration),
                                      int(seed), int(tree_id), tt.h if tt else None, _ptr(zt), int(max_root_moves), _ptr(md),
                                      int(max_turns), float(bias_weight), _ptr(vis), _ptr(rew), _ptr(mov))
    return {"n_root": int(n), "visits": vis, "reward": rew, "moves": mov}


def win_probability(state: np.ndarray, player: int, model: np.ndarray, model_kind: int = 0, max_turns: int = 2500,
                    frontiers: Optional[np.ndarray] = None) -> float:
    """LearnedWinProbabilityEvaluator.predict_player_win_probability for one packed state (logistic or GBT model form)."""
    st = np.ascontiguousarray(np.atleast_1d(state), dtype=STATE_DTYPE)
    md = np.ascontiguousarray(model, dtype=np.float64)
    fr = None if frontiers is None else np.ascontiguousarray(frontiers, dtype=np.uint8)
    return float(lib().orc_win_probability_packed(_ptr(st), _ptr(fr), int(player), _ptr(md), int(model_kind), int(max_turns)))


def mcts_search_shaped(root: np.ndarray, iterations: int, seed: int, model: np.ndarray, tree_id: int = 0, max_turns: int = 2500,
                       root_frontier: Optional[np.ndarray] = None, max_rollout_moves: int = 50, rollout_policy: int = 0,
                       exploration: float = 1.414, shaping_weight: float = 1.0, shaping_gamma: float = 1.0,
                       potential_mode: int = 0, bias_weight: float = 0.0, tt: Optional[OracleTT] = None,
                       zobrist: Optional[np.ndarray] = None, max_root_moves: int = 2048, model_kind: int = 0, tt_namespace: int = 0):
    """UCT with rollouts and potential-based shaping from the learned model (mcts_agent.py:470-554); reward sums are doubles."""
    st = np.ascontiguousarray(np.atleast_1d(root), dtype=STATE_DTYPE)
    md = np.ascontiguousarray(model, dtype=np.float64)
    lib().orc_set_model_kind(int(model_kind))
    lib().orc_set_tt_namespace(int(tt_namespace))
    fr = None if root_frontier is None else np.ascontiguousarray(root_frontier, dtype=np.uint8)
    vis = np.zeros(max_root_moves, dtype=np.int32)
    rew = np.zeros(max_root_moves, dtype=np.float64)
    mov = np.zeros(max_root_moves, dtype=np.uint16)
    zt = None if zobrist is None else np.ascontiguousarray(zobrist, dtype=np.uint64)
    n = lib().orc_mcts_search_shaped(_ptr(st), _ptr(fr), int(iterations), int(max_rollout_moves), int(rollout_policy),
                                     float(exploration), int(seed), int(tree_id), tt.h if tt else None, _ptr(zt),
                                     int(max_root_moves), _ptr(md), int(max_turns), float(bias_weight), float(shaping_weight),
                                     float(shaping_gamma), int(potential_mode), _ptr(vis), _ptr(rew), _ptr(mov))
    return {"n_root": int(n), "visits": vis, "reward": rew, "moves": mov}


def mcts_wide(root: np.ndarray, call_sims: Sequence[int], width: int, seed: int, tree_id: int = 0, max_rollout_moves: int = 50,
              max_depth: int = 0, ucb_variant: int = 0, reward_mode: int = 0, rollout_policy: int = 0, exploration: float = 1.414,
              max_root_moves: int = 2048):
    """Sequential restatement of the round-synchronous search (csrc/mcts_wide.cuh): call_sims = simulations added by each
    search call.  Returns dict(n_root, visits, reward, moves, rounds, sims, nodes)."""
    st = np.ascontiguousarray(np.atleast_1d(root), dtype=STATE_DTYPE)
    calls = np.ascontiguousarray(call_sims, dtype=np.int32)
    vis = np.zeros(max_root_moves, dtype=np.int32)
    rew = np.zeros(max_root_moves, dtype=np.int64)
    mov = np.zeros(max_root_moves, dtype=np.uint16)
    info = np.zeros(3, dtype=np.int32)
    n = lib().orc_mcts_wide(_ptr(st), _ptr(calls), len(calls), int(width), int(max_rollout_moves), int(max_depth), int(ucb_variant),
                            int(reward_mode), int(rollout_policy), float(exploration), int(seed), int(tree_id), int(max_root_moves),
                            _ptr(vis), _ptr(rew), _ptr(mov), _ptr(info))
    return {"n_root": int(n), "visits": vis, "reward": rew, "moves": mov, "rounds": int(info[0]), "sims": int(info[1]),
            "nodes": int(info[2])}


def check_moves_batch(states: np.ndarray, offsets: np.ndarray, ids: np.ndarray, naive: bool = False,
                      threads: int = 0) -> Tuple[int, int]:
    """Compare CSR move lists for the side to move of every state with the oracle's; (mismatches, first bad index)."""
    from concurrent.futures import ThreadPoolExecutor

    st = np.ascontiguousarray(states, dtype=STATE_DTYPE)
    off = np.ascontiguousarray(offsets, dtype=np.uint32)
    mv = np.ascontiguousarray(ids, dtype=np.uint16)
    n = len(st)
    threads = threads or min(32, os.cpu_count() or 1)
    chunk = max(1, -(-n // (threads * 4)))

    def work(lo):
        fb = C.c_int(-1)
        bad = lib().orc_check_moves_batch(_ptr(st), lo, min(n, lo + chunk), _ptr(off), _ptr(mv), int(naive), C.byref(fb))
        return bad, fb.value

    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(work, range(0, n, chunk)))
    bad = sum(r[0] for r in res)
    firsts = [r[1] for r in res if r[1] >= 0]
    return bad, (min(firsts) if firsts else -1)


def check_playouts(seed: int, first_game: int, traces: np.ndarray, game_scores: np.ndarray, policy: int = 0,
                   threads: int = 0) -> Tuple[int, int]:
    """Replay every game with the oracle and compare trace + final scores; (mismatches, first bad index)."""
    from concurrent.futures import ThreadPoolExecutor

    tr = np.ascontiguousarray(traces, dtype=np.uint16)
    gs = np.ascontiguousarray(game_scores, dtype=np.int32)
    n = len(tr)
    threads = threads or min(32, os.cpu_count() or 1)
    chunk = max(1, -(-n // (threads * 8)))

    def work(lo):
        fb = C.c_int(-1)
        bad = lib().orc_check_playouts(int(seed), int(first_game), lo, min(n, lo + chunk), int(policy), _ptr(tr), _ptr(gs),
                                       C.byref(fb))
        return bad, fb.value

    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(work, range(0, n, chunk)))
    firsts = [r[1] for r in res if r[1] >= 0]
    return sum(r[0] for r in res), (min(firsts) if firsts else -1)

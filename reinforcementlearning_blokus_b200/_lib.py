"""ctypes binding of the C ABI declared in include/blokus_b200.h.

The shared library is the product: if it is missing or no CUDA device is present, every compute
call raises.  There is no CPU path behind this module.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# BLOKUS_B200_LIB points at an alternative build of the SAME library (kernel tuning experiments only)
LIB_PATH = os.environ.get("BLOKUS_B200_LIB") or os.path.join(HERE, "libblokus_b200.so")

BK_PASS = 0xFFFF
BK_MAX_TURNS = 96
BK_BITMAP_WORDS = 91 * 20
BK_ZOBRIST_ENTRIES = 2088
STATE_BYTES = 256


class BlokusB200Error(RuntimeError):
    pass


_vp, _i32, _u32, _u64, _i64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_int64

#: name -> (restype, argtypes); kept in one place so tests can check the export list against the header
PROTOTYPES = {
    "bk_init": (_i32, [_i32]),
    "bk_last_error": (C.c_char_p, []),
    "bk_version": (C.c_char_p, []),
    "bk_movegen_count": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "bk_exclusive_scan_u32": (_i32, [_vp, _i32, _vp, _vp]),
    "bk_movegen_emit": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "bk_movegen_fixed": (_i32, [_vp, _vp, _i32, _u32, _vp, _vp, _vp]),
    "bk_legal_moves_host": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "bk_can_place_mask": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "bk_movegen_bitmap": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "bk_compact_rows": (_i32, [_vp, _u32, _vp, _i32, _vp, _vp]),
    "bk_movegen_piece_counts": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "bk_has_moves": (_i32, [_vp, _i32, _vp, _vp]),
    "bk_snapshot_features": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "bk_replay_snapshots": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "bk_is_move_legal": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "bk_step": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "bk_scores": (_i32, [_vp, _i32, _vp, _vp, _vp]),
    "bk_playout": (_i32, [_vp, _i32, _u64, _u64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bk_playout_seats": (_i32, [_vp, _i32, _u64, _u64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bk_zobrist_hash": (_i32, [_vp, _i32, _vp, _vp, _vp]),
    "bk_tt_create": (_i32, [C.POINTER(_vp), _u32]),
    "bk_tt_clear": (_i32, [_vp, _vp]),
    "bk_tt_destroy": (_i32, [_vp]),
    "bk_tt_get": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "bk_tt_put": (_i32, [_vp, _vp, _vp, _i32, _vp]),
    "bk_tt_stats": (_i32, [_vp, _vp, _vp]),
    "bk_tt_evict_oldest": (_i32, [_vp, _u64, _vp]),
    "bk_heuristic_scores": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "bk_heuristic_scores_weighted": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "bk_mcts_workspace_bytes": (_i64, [_vp]),
    "bk_mcts_search": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bk_mcts_wide_workspace_bytes": (_i64, [_vp]),
    "bk_mcts_wide_search": (_i32, [_vp, _i32, _i32, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bk_p2p_window_create": (_i32, [_i32, C.POINTER(_vp), _vp]),
    "bk_p2p_window_open": (_i32, [_vp, C.POINTER(_vp)]),
    "bk_p2p_window_close": (_i32, [_vp]),
    "bk_p2p_window_destroy": (_i32, [_vp]),
    "bk_root_merge_p2p": (_i32, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _u64, _i32, _vp, _vp, _vp]),
}


class MctsParams(C.Structure):
    """bk_mcts_params of include/blokus_b200.h."""

    _fields_ = [
        ("iterations", C.c_int32), ("iterations_total", C.c_int32), ("max_rollout_moves", C.c_int32),
        ("rollout_policy", C.c_int32), ("max_depth", C.c_int32), ("ucb_variant", C.c_int32),
        ("resume", C.c_int32), ("reward_mode", C.c_int32), ("leaf_batch", C.c_int32), ("tt_tree_base", C.c_int32),
        ("exploration", C.c_double),
        ("seed", C.c_uint64), ("tree_offset", C.c_uint64),
        ("leaf_model", C.c_void_p), ("root_frontiers", C.c_void_p), ("leaf_max_turns", C.c_int32), ("leaf_model_kind", C.c_int32),
        ("progressive_bias_weight", C.c_double), ("shaping_weight", C.c_double), ("shaping_gamma", C.c_double),
        ("potential_mode", C.c_int32), ("round_width", C.c_int32),
    ]

_lock = threading.Lock()
_lib = None
_inited_devices = set()


def load() -> C.CDLL:
    """dlopen the in-tree library (no compute, safe without a GPU)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise BlokusB200Error(
                    f"{LIB_PATH} is missing: build it with `python -m reinforcementlearning_blokus_b200.build` "
                    "(there is no CPU fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(code: int) -> None:
    if code != 0:
        lib = _lib if _lib is not None else load()       # never re-enter load()'s lock from an error path
        msg = lib.bk_last_error().decode("utf-8", "replace")
        raise BlokusB200Error(f"blokus_b200 error {code}: {msg}")


def init(device: int = 0) -> C.CDLL:
    """bk_init once per device (the C library keeps one context per device and leaves the current device alone);
    raises BlokusB200Error when the device does not exist or cannot run the kernels."""
    lib = load()
    if device in _inited_devices:
        return lib
    code = lib.bk_init(int(device))          # outside the module lock: check() reads the library's error string
    check(code)
    with _lock:
        _inited_devices.add(device)
    return lib

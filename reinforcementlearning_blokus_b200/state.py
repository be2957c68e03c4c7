"""Packed 256-byte game state (``bk_state`` of include/blokus_b200.h) on the host side."""

from __future__ import annotations

import numpy as np

STATE_BYTES = 256

STATE_DTYPE = np.dtype(
    [
        ("bits", "<u8", (4, 7)),      # int(board.player_bits[p]).to_bytes(56, 'little')  (engine/bitboard.py:13-22)
        ("used", "<u4", (4,)),        # bit piece_id-1                                     (engine/board.py:63)
        ("to_move", "u1"),            # current_player.value - 1                           (engine/board.py:66)
        ("first_mask", "u1"),         # bit p while player_first_move[p]                   (engine/board.py:64)
        ("move_count", "<u2"),        # board.move_count                                   (engine/board.py:67)
        ("reserved", "u1", (12,)),
    ]
)
assert STATE_DTYPE.itemsize == STATE_BYTES


def empty_states(n: int = 1) -> np.ndarray:
    """n start-of-game states: empty board, RED to move, nobody has moved."""
    s = np.zeros(n, dtype=STATE_DTYPE)
    s["first_mask"] = 0xF
    return s


def bits_to_int(words: np.ndarray) -> int:
    """7 little-endian uint64 words -> Python int bitboard (bit = row*20+col)."""
    return int.from_bytes(np.ascontiguousarray(words, dtype="<u8").tobytes(), "little")


def int_to_bits(mask: int) -> np.ndarray:
    return np.frombuffer(int(mask).to_bytes(56, "little"), dtype="<u8").copy()


def grid_of(state) -> np.ndarray:
    """20x20 int grid (0 empty, 1..4 player) of one packed state."""
    g = np.zeros((20, 20), dtype=np.int64)
    for p in range(4):
        m = bits_to_int(state["bits"][p])
        for i in range(400):
            if (m >> i) & 1:
                g[i // 20, i % 20] = p + 1
    return g

"""Piece set and the 91 placement orientations, in the reference's index order.

Host-side source of truth for every table the device code uses (the CUDA header
``csrc/tables_gen.h`` and the straight-line orientation code ``csrc/orient_gen.cuh`` are
emitted from this module by ``tools/gen_tables.py``).

Follows the reference's construction (``engine/pieces.py:160-192`` and ``:367-400``):
for each of the 21 base shapes (``engine/pieces.py:288-356``) the variants are visited in
the order rot0, rot90ccw, rot180, rot270, fliplr, fliplr+rot90, +180, +270, duplicates
are dropped on first occurrence, and cells are expressed as (row, col) offsets from the
bounding-box top-left, sorted ascending.  Two reference quirks are kept on purpose
(SURVEY.md Q1): ids 9 and 10 are the same free tetromino and there is no Z pentomino,
so the set has 88 squares.
"""

from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

BOARD_SIZE = 20
NUM_CELLS = BOARD_SIZE * BOARD_SIZE
NUM_PIECES = 21
NUM_ORIENTATIONS = 91

Cell = Tuple[int, int]

# (id, name, rows) -- rows are strings, '#' = filled.  Same shapes as engine/pieces.py:293-354.
_BASE_SHAPES: Sequence[Tuple[int, str, Sequence[str]]] = (
    (1, "Monomino", ("#",)),
    (2, "Domino", ("##",)),
    (3, "Tromino I", ("###",)),
    (4, "Tromino L", ("#.", "##")),
    (5, "Tetromino I", ("####",)),
    (6, "Tetromino O", ("##", "##")),
    (7, "Tetromino T", ("###", ".#.")),
    (8, "Tetromino L", ("#.", "#.", "##")),
    (9, "Tetromino S", (".##", "##.")),
    (10, "Tetromino Z", ("##.", ".##")),
    (11, "Pentomino F", (".##", "##.", ".#.")),
    (12, "Pentomino I", ("#####",)),
    (13, "Pentomino L", ("#.", "#.", "#.", "##")),
    (14, "Pentomino N", ("#.", "##", ".#", ".#")),
    (15, "Pentomino P", ("##", "##", "#.")),
    (16, "Pentomino T", ("###", ".#.", ".#.")),
    (17, "Pentomino U", ("#.#", "###")),
    (18, "Pentomino V", ("#..", "#..", "###")),
    (19, "Pentomino W", ("#..", "##.", ".##")),
    (20, "Pentomino X", (".#.", "###", ".#.")),
    (21, "Pentomino Y", ("#.", "##", "#.", "#.")),
)


def _to_matrix(rows: Sequence[str]) -> List[List[int]]:
    return [[1 if ch == "#" else 0 for ch in row] for row in rows]


def _rot90_ccw(m: List[List[int]]) -> List[List[int]]:
    """Counter-clockwise quarter turn (what numpy.rot90 does with k=1)."""
    h, w = len(m), len(m[0])
    return [[m[j][w - 1 - i] for j in range(h)] for i in range(w)]


def _fliplr(m: List[List[int]]) -> List[List[int]]:
    return [list(reversed(row)) for row in m]


def _cells(m: List[List[int]]) -> Tuple[Cell, ...]:
    pts = [(r, c) for r, row in enumerate(m) for c, v in enumerate(row) if v]
    r0 = min(r for r, _ in pts)
    c0 = min(c for _, c in pts)
    return tuple(sorted((r - r0, c - c0) for r, c in pts))


class Orientation:
    """One placement shape: global index, owning piece, per-piece index and its cells."""

    __slots__ = ("gidx", "piece_id", "index", "cells", "height", "width")

    def __init__(self, gidx: int, piece_id: int, index: int, cells: Tuple[Cell, ...]):
        self.gidx = gidx
        self.piece_id = piece_id
        self.index = index
        self.cells = cells
        self.height = 1 + max(r for r, _ in cells)
        self.width = 1 + max(c for _, c in cells)

    def matrix(self) -> List[List[int]]:
        m = [[0] * self.width for _ in range(self.height)]
        for r, c in self.cells:
            m[r][c] = 1
        return m

    def __repr__(self) -> str:  # pragma: no cover - debugging aid
        return f"Orientation(g={self.gidx}, piece={self.piece_id}, o={self.index}, cells={self.cells})"


def _build() -> Tuple[Dict[int, List[Orientation]], List[Orientation], Dict[int, int], Dict[int, str]]:
    per_piece: Dict[int, List[Orientation]] = {}
    flat: List[Orientation] = []
    sizes: Dict[int, int] = {}
    names: Dict[int, str] = {}
    for pid, name, rows in _BASE_SHAPES:
        base = _to_matrix(rows)
        variants = []
        m = base
        for _ in range(4):
            variants.append(m)
            m = _rot90_ccw(m)
        m = _fliplr(base)
        for _ in range(4):
            variants.append(m)
            m = _rot90_ccw(m)
        seen = []
        for v in variants:
            key = _cells(v)
            if key not in seen:
                seen.append(key)
        per_piece[pid] = []
        for k, key in enumerate(seen):
            o = Orientation(len(flat), pid, k, key)
            per_piece[pid].append(o)
            flat.append(o)
        sizes[pid] = len(seen[0])
        names[pid] = name
    return per_piece, flat, sizes, names


PIECE_ORIENTATIONS, ORIENTATIONS, PIECE_SIZES, PIECE_NAMES = _build()
assert len(ORIENTATIONS) == NUM_ORIENTATIONS and len(PIECE_ORIENTATIONS) == NUM_PIECES

#: first global orientation index of each piece id (index 0 unused), plus a sentinel at 22
PIECE_FIRST_ORIENTATION: List[int] = [0] * (NUM_PIECES + 2)
for _pid in range(1, NUM_PIECES + 1):
    PIECE_FIRST_ORIENTATION[_pid] = PIECE_ORIENTATIONS[_pid][0].gidx
PIECE_FIRST_ORIENTATION[NUM_PIECES + 1] = NUM_ORIENTATIONS

MAX_MOVE_ID = NUM_ORIENTATIONS * NUM_CELLS  # exclusive upper bound (36 400)
PASS_ID = 0xFFFF


def encode_move(piece_id: int, orientation: int, row: int, col: int) -> int:
    """(piece, per-piece orientation, bbox row, bbox col) -> canonical 16-bit move id."""
    g = PIECE_ORIENTATIONS[piece_id][orientation].gidx
    return g * NUM_CELLS + row * BOARD_SIZE + col


def decode_move(move_id: int) -> Tuple[int, int, int, int]:
    """Canonical move id -> (piece, per-piece orientation, bbox row, bbox col)."""
    g, cell = divmod(int(move_id), NUM_CELLS)
    o = ORIENTATIONS[g]
    return o.piece_id, o.index, cell // BOARD_SIZE, cell % BOARD_SIZE

"""``LegalMoveGenerator`` / ``Move`` with the reference's interface, generated on the GPU.

Interface followed: ``engine/move_generator.py:90-106`` (Move), ``:109-151`` (generator construction and
``get_legal_moves``), ``:872-959`` (per-piece moves, ``is_move_legal``, counts), ``:961-975``
(``has_legal_moves``) and ``:1220-1233`` (shared instance).  The reference's three generator variants
(naive / frontier+grid / frontier+bitboard, selected by environment flags, ``:43-87``) collapse into one
device kernel; there are no flags and no CPU path.

Moves are returned in ascending canonical order (piece, orientation, anchor row, anchor col) -- the order
of the reference's naive generator.  The reference's default generator returns the same *set* in an
order that depends on CPython set iteration.
"""

from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from . import pieces as P
from . import single
from .board import Board, Player, Position


class Move:
    """(piece_id 1..21, orientation index, bbox-top-left anchor row, col) -- engine/move_generator.py:90."""

    __slots__ = ("piece_id", "orientation", "anchor_row", "anchor_col")

    def __init__(self, piece_id: int, orientation: int, anchor_row: int, anchor_col: int):
        self.piece_id = piece_id
        self.orientation = orientation
        self.anchor_row = anchor_row
        self.anchor_col = anchor_col

    def get_positions(self, piece_orientations: List) -> List[Position]:
        shape = piece_orientations[self.orientation]
        rows, cols = shape.shape
        return [Position(self.anchor_row + i, self.anchor_col + j)
                for i in range(rows) for j in range(cols) if shape[i, j] == 1]

    # -- canonical 16-bit id -----------------------------------------------------------------------
    def to_id(self) -> Optional[int]:
        """Canonical move id, or None when the fields do not name a placement inside the board."""
        per_piece = P.PIECE_ORIENTATIONS.get(self.piece_id)
        if per_piece is None or not 0 <= self.orientation < len(per_piece):
            return None
        o = per_piece[self.orientation]
        if not (0 <= self.anchor_row <= P.BOARD_SIZE - o.height and 0 <= self.anchor_col <= P.BOARD_SIZE - o.width):
            return None
        return o.gidx * P.NUM_CELLS + self.anchor_row * P.BOARD_SIZE + self.anchor_col

    @classmethod
    def from_id(cls, move_id: int) -> "Move":
        return cls(*P.decode_move(move_id))

    def __eq__(self, other):
        return isinstance(other, Move) and (self.piece_id, self.orientation, self.anchor_row, self.anchor_col) == (
            other.piece_id, other.orientation, other.anchor_row, other.anchor_col)

    def __hash__(self):
        return hash((self.piece_id, self.orientation, self.anchor_row, self.anchor_col))

    def __str__(self):
        return (f"Move(piece_id={self.piece_id}, orientation={self.orientation}, "
                f"anchor=({self.anchor_row}, {self.anchor_col}))")

    __repr__ = __str__


_ORIENT_PIECE = np.array([o.piece_id for o in P.ORIENTATIONS], dtype=np.int64)
_ORIENT_INDEX = np.array([o.index for o in P.ORIENTATIONS], dtype=np.int64)


class Piece:
    """id / name / shape / size record (engine/pieces.py:13-27)."""

    def __init__(self, pid: int):
        self.id = pid
        self.name = P.PIECE_NAMES[pid]
        self.shape = np.array(P.PIECE_ORIENTATIONS[pid][0].matrix(), dtype=int)
        self.size = P.PIECE_SIZES[pid]


class MoveList(list):
    """The list get_legal_moves returns: a plain list of Move objects that also remembers that it is in ascending
    canonical id order (``ascending_ids``: the ids as a numpy array), so that agents handed this very list do not have to
    recompute and sort 200+ ids per call.  Any copy or slice is an ordinary list again."""

    ascending_ids: Optional[np.ndarray] = None

    def _forget(name):                     # any in-place change makes it an ordinary list again
        def method(self, *a, **k):
            self.ascending_ids = None
            return getattr(list, name)(self, *a, **k)
        method.__name__ = name
        return method

    for _n in ("__setitem__", "__delitem__", "__iadd__", "__imul__", "append", "extend", "insert", "pop", "remove", "clear",
               "sort", "reverse"):
        locals()[_n] = _forget(_n)
    del _n, _forget


class LegalMoveGenerator:
    """Drop-in for the reference generator; stateless apart from the immutable shape caches."""

    def __init__(self):
        self.all_pieces = [Piece(pid) for pid in range(1, P.NUM_PIECES + 1)]
        # read directly by callers (engine/game.py:92-99, mcts/mcts_agent.py:149-162, agents/heuristic_agent.py:84)
        self.piece_orientations_cache: Dict[int, List[np.ndarray]] = {
            pid: [np.array(o.matrix(), dtype=int) for o in P.PIECE_ORIENTATIONS[pid]] for pid in P.PIECE_ORIENTATIONS}
        self.piece_position_cache: Dict[int, List[List[tuple]]] = {
            pid: [list(o.cells) for o in P.PIECE_ORIENTATIONS[pid]] for pid in P.PIECE_ORIENTATIONS}

    # ---- generation ---------------------------------------------------------------------------------
    def get_legal_move_ids(self, board: Board, player: Player) -> np.ndarray:
        return single.legal_move_ids(board.to_packed(), player.value - 1)

    def get_legal_moves(self, board: Board, player: Player) -> List[Move]:
        ids = self.get_legal_move_ids(board, player).astype(np.int64)
        g, cell = np.divmod(ids, P.NUM_CELLS)
        pid, ori = _ORIENT_PIECE[g].tolist(), _ORIENT_INDEX[g].tolist()
        row, col = np.divmod(cell, P.BOARD_SIZE)
        out = MoveList(Move(a, b, c, d) for a, b, c, d in zip(pid, ori, row.tolist(), col.tolist()))
        out.ascending_ids = ids
        return out

    # the reference exposes its variants under these names; tests and benchmarks call them directly
    _get_legal_moves_naive = get_legal_moves
    _get_legal_moves_frontier = get_legal_moves

    def get_legal_moves_for_piece(self, board: Board, player: Player, piece_id: int) -> List[Move]:
        if piece_id in board.player_pieces_used[player] or piece_id not in P.PIECE_ORIENTATIONS:
            return []
        lo = P.PIECE_FIRST_ORIENTATION[piece_id] * P.NUM_CELLS
        hi = P.PIECE_FIRST_ORIENTATION[piece_id + 1] * P.NUM_CELLS
        return [Move.from_id(i) for i in self.get_legal_move_ids(board, player) if lo <= i < hi]

    def is_move_legal(self, board: Board, player: Player, move: Move) -> bool:
        """False (never an exception) for spent pieces, bad indices and off-board anchors
        (engine/move_generator.py:912-946; tests/test_engine.py:260-274 uses Move(1, 0, -1, -1))."""
        if move.piece_id in board.player_pieces_used[player]:
            return False
        mid = move.to_id()
        if mid is None:
            return False
        return single.is_legal(board.to_packed(), player.value - 1, mid)

    def get_move_count(self, board: Board, player: Player) -> int:
        return single.count_moves(board.to_packed(), player.value - 1)

    def has_legal_moves(self, board: Board, player: Player) -> bool:
        return bool((single.has_moves_mask(board.to_packed()) >> (player.value - 1)) & 1)

    def get_game_state_summary(self, board: Board) -> dict:
        mask = single.has_moves_mask(board.to_packed())
        return {
            "current_player": board.current_player.name,
            "move_count": board.move_count,
            "game_over": board.game_over,
            "can_move": {p.name: bool((mask >> (p.value - 1)) & 1) for p in Player},
            "pieces_used": {p.name: len(board.player_pieces_used[p]) for p in Player},
        }


_SHARED_GENERATOR: Optional[LegalMoveGenerator] = None


def get_shared_generator() -> LegalMoveGenerator:
    """Process-wide generator instance (engine/move_generator.py:1220-1233)."""
    global _SHARED_GENERATOR
    if _SHARED_GENERATOR is None:
        _SHARED_GENERATOR = LegalMoveGenerator()
    return _SHARED_GENERATOR

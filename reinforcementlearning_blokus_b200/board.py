"""Host-side ``Board`` with the attribute and method surface of the reference's ``engine.board.Board``.

The board object is a plain state holder (grid, 400-bit masks, spent pieces, side to move); every
rule decision -- ``can_place_piece`` and, through ``move_generator``/``game``, legal-move generation and
game-over detection -- runs on the GPU.  Interface followed: ``engine/board.py:15-27`` (Player),
``:29-40`` (Position), ``:43-78`` (state), ``:136-220`` (placement rules), ``:515-560`` (place / turn),
``:562-591`` (score / winner), ``:643-660`` (copy).

``get_frontier`` follows the reference's bookkeeping (``engine/board.py:315-406``): each player's set is updated
only by that player's own placements, so it can keep cells an opponent has taken since.  The sets
are 400-bit masks here; they feed the snapshot features (``winprob.py``), not the rules -- legality on the device is
computed from the cells.  ``_compute_full_frontier`` is the exact view.  A board restored from a packed state starts
from the exact sets (plus the start corner before the first move), like a board rebuilt from the wire.
"""

from __future__ import annotations

import struct
from dataclasses import dataclass
from enum import Enum
from typing import Dict, Iterable, List, Optional, Set, Tuple

import numpy as np

from .state import STATE_DTYPE

SIZE = 20
_META = struct.Struct("<4IBBH")        # used[4], to_move, first_mask, move_count of bk_state
_PAD = bytes(12)
_FULL = (1 << 400) - 1
_COL0 = sum(1 << (r * SIZE) for r in range(SIZE))
_COL19 = _COL0 << (SIZE - 1)


class Player(Enum):
    """Turn order RED -> BLUE -> YELLOW -> GREEN (engine/board.py:15-24)."""

    RED = 1
    BLUE = 2
    YELLOW = 3
    GREEN = 4


_PLAYERS = list(Player)


@dataclass
class Position:
    row: int
    col: int

    def __hash__(self):
        return hash((self.row, self.col))

    def __eq__(self, other):
        return self.row == other.row and self.col == other.col


def _cells_mask(cells: Iterable[Tuple[int, int]]) -> int:
    m = 0
    for r, c in cells:
        m |= 1 << (r * SIZE + c)
    return m


def _neighbours(mask: int) -> Tuple[int, int]:
    """(orthogonal, diagonal) neighbour masks of a 400-bit cell mask."""
    east = (mask & ~_COL19) << 1
    west = (mask & ~_COL0) >> 1
    horiz = east | west
    orth = (horiz | (mask << SIZE) | (mask >> SIZE)) & _FULL
    diag = ((horiz << SIZE) | (horiz >> SIZE)) & _FULL
    return orth, diag


class Board:
    """20x20 Blokus board; 0 = empty, 1..4 = ``Player.value``."""

    SIZE = SIZE

    def __init__(self):
        self.grid = np.zeros((SIZE, SIZE), dtype=int)
        self.player_start_corners: Dict[Player, Position] = {
            Player.RED: Position(0, 0),
            Player.BLUE: Position(0, SIZE - 1),
            Player.YELLOW: Position(SIZE - 1, SIZE - 1),
            Player.GREEN: Position(SIZE - 1, 0),
        }
        self.player_pieces_used: Dict[Player, Set[int]] = {p: set() for p in Player}
        self.player_first_move: Dict[Player, bool] = dict.fromkeys(Player, True)
        self.game_over = False
        self.current_player = Player.RED
        self.move_count = 0
        self.occupied_bits = 0
        self.player_bits: Dict[Player, int] = {p: 0 for p in Player}
        # engine/board.py:369-406: every frontier starts as the player's start corner
        self.player_frontier_bits: Dict[Player, int] = {
            p: 1 << (c.row * SIZE + c.col) for p, c in self.player_start_corners.items()}

    # ---- packed state --------------------------------------------------------------------------
    def to_packed(self) -> np.ndarray:
        """This board as one ``bk_state`` record (numpy scalar of STATE_DTYPE), assembled as raw bytes."""
        used, mask = [0, 0, 0, 0], 0
        for k, p in enumerate(_PLAYERS):
            u = 0
            for pid in self.player_pieces_used[p]:
                if 1 <= pid <= 21:
                    u |= 1 << (pid - 1)
            used[k] = u
            if self.player_first_move[p]:
                mask |= 1 << k
        bits = self.player_bits
        buf = bytearray(b"".join(int(bits[p]).to_bytes(56, "little") for p in _PLAYERS))
        buf += _META.pack(used[0], used[1], used[2], used[3], self.current_player.value - 1, mask, self.move_count & 0xFFFF)
        buf += _PAD
        return np.frombuffer(buf, dtype=STATE_DTYPE)[0]

    @classmethod
    def from_packed(cls, state) -> "Board":
        """Inverse of to_packed (the "state from wire" recipe of engine-service/app.py:27-50)."""
        b = cls()
        for k, p in enumerate(_PLAYERS):
            m = int.from_bytes(np.ascontiguousarray(state["bits"][k], dtype="<u8").tobytes(), "little")
            b.player_bits[p] = m
            b.occupied_bits |= m
            for i in range(SIZE * SIZE):
                if (m >> i) & 1:
                    b.grid[i // SIZE, i % SIZE] = p.value
            used = int(state["used"][k])
            b.player_pieces_used[p] = {pid for pid in range(1, 22) if (used >> (pid - 1)) & 1}
            b.player_first_move[p] = bool((int(state["first_mask"]) >> k) & 1)
        b.current_player = _PLAYERS[int(state["to_move"]) & 3]
        b.move_count = int(state["move_count"])
        b.rebuild_frontiers()
        return b

    def rebuild_frontiers(self) -> None:
        """Frontier sets of a board restored from its cells (engine-service/app.py:37): the exact set for a player who
        has moved, the start corner for one who has not."""
        for p in _PLAYERS:
            orth, diag = _neighbours(self.player_bits[p])
            m = diag & ~orth & ~self.occupied_bits & _FULL
            c = self.player_start_corners[p]
            if self.player_first_move[p] and self.grid[c.row, c.col] == 0:
                m |= 1 << (c.row * SIZE + c.col)
            self.player_frontier_bits[p] = m

    # ---- cell helpers (engine/board.py:80-134) ---------------------------------------------------
    def is_valid_position(self, pos: Position) -> bool:
        return 0 <= pos.row < SIZE and 0 <= pos.col < SIZE

    def get_cell(self, pos: Position) -> int:
        if not self.is_valid_position(pos):
            return -1
        return self.grid[pos.row, pos.col]

    def set_cell(self, pos: Position, value: int) -> None:
        if self.is_valid_position(pos):
            self.grid[pos.row, pos.col] = value

    def is_empty(self, pos: Position) -> bool:
        return self.get_cell(pos) == 0

    def get_player_at(self, pos: Position) -> Optional[Player]:
        v = self.get_cell(pos)
        return None if v <= 0 else Player(v)

    def _around(self, pos: Position, deltas) -> List[Position]:
        out = []
        for dr, dc in deltas:
            q = Position(pos.row + dr, pos.col + dc)
            if self.is_valid_position(q):
                out.append(q)
        return out

    def get_adjacent_positions(self, pos: Position) -> List[Position]:
        return self._around(pos, [(dr, dc) for dr in (-1, 0, 1) for dc in (-1, 0, 1) if dr or dc])

    def get_edge_adjacent_positions(self, pos: Position) -> List[Position]:
        return self._around(pos, [(-1, 0), (1, 0), (0, -1), (0, 1)])

    def get_corner_adjacent_positions(self, pos: Position) -> List[Position]:
        return self._around(pos, [(-1, -1), (-1, 1), (1, -1), (1, 1)])

    # ---- rules: decided on the device ---------------------------------------------------------------
    def can_place_piece(self, piece_positions: List[Position], player: Player) -> bool:
        """engine/board.py:136-220.  Off-board or empty position lists are rejected here (they have no
        mask representation); everything else is one ``bk_can_place_mask`` launch."""
        if not piece_positions:
            return False
        for pos in piece_positions:
            if not self.is_valid_position(pos):
                return False
        from . import single

        return single.can_place(self.to_packed(), player.value - 1, _cells_mask((p.row, p.col) for p in piece_positions))

    def place_piece(self, piece_positions: List[Position], player: Player, piece_id: int, validate: bool = True) -> bool:
        """engine/board.py:515-555: colour the cells, spend the piece, clear the first-move flag, count the
        move and advance the turn from ``current_player``."""
        if validate and not self.can_place_piece(piece_positions, player):
            return False
        cells = [(pos.row, pos.col) for pos in piece_positions]
        for r, c in cells:
            self.grid[r, c] = player.value
        m = _cells_mask(cells)
        self.occupied_bits |= m
        self.player_bits[player] |= m
        self.player_pieces_used[player].add(piece_id)
        self.player_first_move[player] = False
        # engine/board.py:315-367: the placed cells and their edge neighbours leave the mover's set; empty corner
        # neighbours that touch no own cell by an edge join it
        gone_orth, join_diag = _neighbours(m)
        own_orth, _ = _neighbours(self.player_bits[player])
        self.player_frontier_bits[player] = ((self.player_frontier_bits[player] & ~(m | gone_orth))
                                             | (join_diag & ~self.occupied_bits & ~own_orth)) & _FULL
        self.move_count += 1
        self._update_current_player()
        return True

    def _update_current_player(self) -> None:
        self.current_player = _PLAYERS[(_PLAYERS.index(self.current_player) + 1) % len(_PLAYERS)]

    # ---- frontier -----------------------------------------------------------------------------------
    def get_frontier(self, player: Player) -> Set[Tuple[int, int]]:
        """The incrementally maintained set of engine/board.py:247-259 (see the module docstring)."""
        m = self.player_frontier_bits[player]
        return {(i // SIZE, i % SIZE) for i in range(SIZE * SIZE) if (m >> i) & 1}

    def frontier_words(self) -> np.ndarray:
        """int64[4, 7]: the four frontier sets as bitboards, the `frontiers` argument of bk_snapshot_features."""
        out = np.zeros((4, 7), dtype="<i8")
        for k, p in enumerate(_PLAYERS):
            out[k] = np.frombuffer(int(self.player_frontier_bits[p]).to_bytes(56, "little"), dtype="<i8")
        return out

    def _compute_full_frontier(self, player: Player) -> Set[Tuple[int, int]]:
        own = self.player_bits[player]
        orth, diag = _neighbours(own)
        m = diag & ~orth & ~self.occupied_bits & _FULL
        return {(i // SIZE, i % SIZE) for i in range(SIZE * SIZE) if (m >> i) & 1}

    # ---- score / end of game --------------------------------------------------------------------------
    def get_score(self, player: Player) -> int:
        """Cells covered + 15 when all 21 pieces are spent (engine/board.py:562-577); device K3."""
        from . import single

        return single.scores(self.to_packed())[1][player.value - 1]

    def get_winner(self) -> Optional[Player]:
        if not self.game_over:
            return None
        from . import single

        board_scores = single.scores(self.to_packed())[1]
        best = max(range(4), key=lambda k: board_scores[k])      # first maximum, as max(dict, key=...) does
        return _PLAYERS[best]

    def is_game_over(self) -> bool:
        return self.game_over

    def assert_bitboard_consistent(self) -> None:
        """Masks and grid describe the same position (engine/board.py:593-641)."""
        occ = 0
        for p in Player:
            m = self.player_bits[p]
            assert m & occ == 0, "player masks overlap"
            occ |= m
            for i in range(SIZE * SIZE):
                assert ((m >> i) & 1) == int(self.grid[i // SIZE, i % SIZE] == p.value), f"mask/grid mismatch at {divmod(i, SIZE)}"
        assert occ == self.occupied_bits, "occupied_bits out of sync"

    def copy(self) -> "Board":
        b = object.__new__(Board)
        b.grid = self.grid.copy()
        b.player_start_corners = self.player_start_corners
        b.player_pieces_used = {k: set(v) for k, v in self.player_pieces_used.items()}
        b.player_first_move = dict(self.player_first_move)
        b.game_over = self.game_over
        b.current_player = self.current_player
        b.move_count = self.move_count
        b.occupied_bits = self.occupied_bits
        b.player_bits = dict(self.player_bits)
        b.player_frontier_bits = dict(self.player_frontier_bits)
        return b

    def __str__(self) -> str:
        return "\n".join("".join("." if v == 0 else str(int(v)) for v in row) for row in self.grid)

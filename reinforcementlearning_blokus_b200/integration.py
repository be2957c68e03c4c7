"""Plugging the device path into a LIVE copy of the reference (when its ``engine`` package is importable).

``install()`` replaces the reference's shared generator (``engine.move_generator._SHARED_GENERATOR``,
engine/move_generator.py:1220-1233) by an adapter whose ``get_legal_moves`` / ``is_move_legal`` /
``has_legal_moves`` read the reference's own ``Board`` object, run on the GPU and hand back the reference's own
``Move`` objects -- so ``BlokusGame``, every agent constructor and ``MCTSNode`` (they all call
``get_shared_generator()``) pick it up without further changes.  ``uninstall()`` restores the original.

Nothing here is needed on a machine without the reference; see INTEGRATION.md for the maintainer-side stub.
"""

from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import pieces as P
from . import single
from .state import STATE_DTYPE

_original = None


def pack_reference_board(board) -> np.ndarray:
    """reference ``engine.board.Board`` -> one packed ``bk_state`` (same recipe as engine-service/app.py:27-50)."""
    from engine.board import Player  # the reference's enum

    s = np.zeros((), dtype=STATE_DTYPE)
    mask = 0
    for k, p in enumerate(Player):
        s["bits"][k] = np.frombuffer(int(board.player_bits[p]).to_bytes(56, "little"), dtype="<u8")
        used = 0
        for pid in board.player_pieces_used[p]:
            if 1 <= pid <= 21:
                used |= 1 << (pid - 1)
        s["used"][k] = used
        if board.player_first_move[p]:
            mask |= 1 << k
    s["first_mask"] = mask
    s["to_move"] = board.current_player.value - 1
    s["move_count"] = int(board.move_count) & 0xFFFF
    return s


def adapter_class():
    """A subclass of the reference's LegalMoveGenerator whose queries run on the device: the public three
    (``get_legal_moves`` / ``is_move_legal`` / ``has_legal_moves``, engine/move_generator.py:130, 912, 961), the
    frontier + bitboard generator they dispatch to (``_get_legal_moves_frontier`` :261, ``_has_any_legal_move_frontier``
    :977, ``get_move_count`` :948) and the coordinate legality test (``is_placement_legal_bitboard_coords`` :682).
    ``_get_legal_moves_naive`` stays the reference's: it is the trusted oracle of the reference's own tests."""
    from engine.move_generator import LegalMoveGenerator as RefGenerator
    from engine.move_generator import Move as RefMove

    base = getattr(RefGenerator, "_b200_base", RefGenerator)       # never stack adapters on adapters

    class B200LegalMoveGenerator(base):
        _b200_base = base

        def get_legal_moves(self, board, player) -> List:
            ids = single.legal_move_ids(pack_reference_board(board), player.value - 1)
            return [RefMove(*P.decode_move(int(i))) for i in ids]

        _get_legal_moves_frontier = get_legal_moves

        def get_move_count(self, board, player) -> int:
            return single.count_moves(pack_reference_board(board), player.value - 1)

        def is_move_legal(self, board, player, move) -> bool:
            if move.piece_id in board.player_pieces_used[player]:
                return False
            per_piece = P.PIECE_ORIENTATIONS.get(move.piece_id)
            if per_piece is None or not 0 <= move.orientation < len(per_piece):
                return False
            o = per_piece[move.orientation]
            if not (0 <= move.anchor_row <= 20 - o.height and 0 <= move.anchor_col <= 20 - o.width):
                return False
            mid = o.gidx * 400 + move.anchor_row * 20 + move.anchor_col
            return single.is_legal(pack_reference_board(board), player.value - 1, mid)

        def has_legal_moves(self, board, player) -> bool:
            return bool((single.has_moves_mask(pack_reference_board(board)) >> (player.value - 1)) & 1)

        _has_any_legal_move_frontier = has_legal_moves

        def is_placement_legal_bitboard_coords(self, board, player, placement_coords, *, is_first_move: bool = False) -> bool:
            mask = _cells_to_mask(placement_coords)
            if mask is None:
                return False
            s = pack_reference_board(board)
            k = player.value - 1                        # the caller states whether this is a first move (:688)
            s["first_mask"] = (int(s["first_mask"]) & ~(1 << k)) | (int(bool(is_first_move)) << k)
            return single.can_place(s, k, mask)

    return B200LegalMoveGenerator


def make_adapter():
    """One device-backed generator object (see adapter_class)."""
    return adapter_class()()


def _cells_to_mask(cells) -> Optional[int]:
    """(row, col) pairs -> 400-bit mask, bit row*20+col; None when a cell is off the board or there is none."""
    mask = 0
    for r, c in cells:
        if not (0 <= r < 20 and 0 <= c < 20):
            return None
        mask |= 1 << (r * 20 + c)
    return mask or None


def _device_can_place_piece(self, piece_positions, player) -> bool:
    """Board.can_place_piece (engine/board.py:136-172) on the device: bk_can_place_mask."""
    mask = _cells_to_mask((p.row, p.col) for p in piece_positions) if piece_positions else None
    if mask is None:
        return False
    return single.can_place(pack_reference_board(self), player.value - 1, mask)


def install(replace_class: bool = False, patch_board: bool = False) -> Optional[object]:
    """Swap the reference's shared generator for the device-backed adapter; returns the adapter.

    replace_class: also rebind ``engine.move_generator.LegalMoveGenerator`` (and every loaded module attribute that
    holds the class), so code that builds its own ``LegalMoveGenerator()`` -- the reference's tests and its benchmark
    do -- gets the device generator too.  patch_board: ``Board.can_place_piece`` runs on the device as well."""
    global _original
    import sys

    import engine.move_generator as mg

    if _original is None:
        _original = {"shared": mg._SHARED_GENERATOR, "class": getattr(mg.LegalMoveGenerator, "_b200_base", mg.LegalMoveGenerator),
                     "rebound": [], "can_place": None}
    cls = adapter_class()
    mg._SHARED_GENERATOR = cls()
    if replace_class:
        old = _original["class"]
        for mod in list(sys.modules.values()):
            for attr, val in list(vars(mod).items()) if mod is not None and hasattr(mod, "__dict__") else ():
                if val is old and attr != "_b200_base":
                    _original["rebound"].append((mod, attr))
                    setattr(mod, attr, cls)
    if patch_board:
        import engine.board as eb

        if _original["can_place"] is None:
            _original["can_place"] = eb.Board.can_place_piece
        eb.Board.can_place_piece = _device_can_place_piece
    return mg._SHARED_GENERATOR


def uninstall() -> None:
    global _original
    import engine.move_generator as mg

    if _original is None:
        return
    mg._SHARED_GENERATOR = _original["shared"]
    for mod, attr in _original["rebound"]:
        setattr(mod, attr, _original["class"])
    if _original["can_place"] is not None:
        import engine.board as eb

        eb.Board.can_place_piece = _original["can_place"]
    _original = None


# ---- the search agents of a live reference ---------------------------------------------------------------------------
_patched = []        # (module, attribute, original object)


def _to_host_board(board):
    from .board import Board

    return Board.from_packed(pack_reference_board(board))


def make_agent_adapters():
    """(MCTSAgent, FastMCTSAgent) classes for a live reference: same constructors as mcts/mcts_agent.py:194 and
    agents/fast_mcts_agent.py:89, taking and returning the REFERENCE's Board / Player / Move objects; the search runs in
    this package's device agents.  (A board handed over this way carries the exact frontier sets, like one rebuilt
    from the wire; only the snapshot features of the learned evaluator look at them.)"""
    from engine.move_generator import Move as RefMove

    from . import agents as _agents
    from . import mcts as _mcts
    from .board import Player as HostPlayer
    from .move_generator import Move as HostMove

    def host_moves(legal_moves):
        return [HostMove(m.piece_id, m.orientation, m.anchor_row, m.anchor_col) for m in legal_moves]

    def ref_move(m):
        return None if m is None else RefMove(m.piece_id, m.orientation, m.anchor_row, m.anchor_col)

    class B200MCTSAgent(_mcts.MCTSAgent):
        def select_action(self, board, player, legal_moves):
            return ref_move(super().select_action(_to_host_board(board), HostPlayer(player.value), host_moves(legal_moves)))

    class B200FastMCTSAgent(_agents.FastMCTSAgent):
        def think(self, board, player, legal_moves, time_budget_ms):
            out = super().think(_to_host_board(board), HostPlayer(player.value), host_moves(legal_moves), time_budget_ms)
            return dict(out, move=ref_move(out.get("move")))

    return B200MCTSAgent, B200FastMCTSAgent


def install_agents():
    """Replace MCTSAgent / FastMCTSAgent in every loaded module of the reference that holds them (their defining
    modules, analytics.tournament.arena_runner, webapi ... import them by name) with the device-backed adapters.
    Returns the two adapter classes; uninstall_agents() restores the originals."""
    import sys

    import agents.fast_mcts_agent as ref_fast
    import mcts.mcts_agent as ref_mcts

    new_mcts, new_fast = make_agent_adapters()
    for old, new in ((ref_mcts.MCTSAgent, new_mcts), (ref_fast.FastMCTSAgent, new_fast)):
        for mod in list(sys.modules.values()):
            if mod is None or not getattr(mod, "__name__", "").split(".")[0] in {"mcts", "agents", "analytics", "webapi", "engine", "scripts"}:
                continue
            for attr, val in list(vars(mod).items()):
                if val is old:
                    _patched.append((mod, attr, old))
                    setattr(mod, attr, new)
    return new_mcts, new_fast


def uninstall_agents() -> None:
    while _patched:
        mod, attr, old = _patched.pop()
        setattr(mod, attr, old)

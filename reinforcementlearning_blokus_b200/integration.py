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


def make_adapter():
    """A subclass of the reference's LegalMoveGenerator whose three public queries run on the device."""
    from engine.move_generator import LegalMoveGenerator as RefGenerator
    from engine.move_generator import Move as RefMove

    class B200LegalMoveGenerator(RefGenerator):
        def get_legal_moves(self, board, player) -> List:
            ids = single.legal_move_ids(pack_reference_board(board), player.value - 1)
            return [RefMove(*P.decode_move(int(i))) for i in ids]

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

    return B200LegalMoveGenerator()


def install() -> Optional[object]:
    """Swap the reference's shared generator for the device-backed adapter; returns the adapter."""
    global _original
    import engine.move_generator as mg

    if _original is None:
        _original = mg._SHARED_GENERATOR
    mg._SHARED_GENERATOR = make_adapter()
    return mg._SHARED_GENERATOR


def uninstall() -> None:
    global _original
    import engine.move_generator as mg

    mg._SHARED_GENERATOR = _original
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

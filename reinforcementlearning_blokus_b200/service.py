"""The stateless ``POST /think`` contract of the reference's engine service (engine-service/app.py:13-72),
served by the device agent.  ``think(payload)`` takes the request body as a dict and returns the response
dict; wiring it to an HTTP framework is one line and is deliberately not done here
(web transport is out of scope).

    payload = {"gameState": {"board": [[..20x20..]], "pieces_used": {"RED": [ids], ...},
                             "current_player": "RED", "move_count": 12},
               "legalMoves": [{"piece_id", "orientation", "anchor_row", "anchor_col"}, ...],
               "timeBudgetMs": 200}
    -> {"move": {...} | None, "stats": {...FastMCTSAgent stats schema...}}
"""

from __future__ import annotations

from typing import Any, Dict, List

import numpy as np

from .agents import FastMCTSAgent
from .board import Board, Player
from .move_generator import Move


def rebuild_board(game_state: Dict[str, Any]) -> Board:
    """State from wire: grid + pieces_used + current_player + move_count; first-move flags follow from
    pieces_used (engine-service/app.py:27-50)."""
    board = Board()
    grid = np.array(game_state.get("board", board.grid.tolist()), dtype=int)
    if grid.shape != (20, 20):
        raise ValueError("board must be 20x20")
    board.grid = grid
    board.move_count = int(game_state.get("move_count", 0))
    board.current_player = Player[game_state.get("current_player", "RED")]
    pieces_used = game_state.get("pieces_used", {})
    for player in Player:
        used = set(pieces_used.get(player.name, []))
        board.player_pieces_used[player] = used
        board.player_first_move[player] = len(used) == 0
    board.occupied_bits = 0
    board.player_bits = {p: 0 for p in Player}
    for r in range(20):
        for c in range(20):
            v = int(grid[r, c])
            if v > 0:
                bit = 1 << (r * 20 + c)
                board.occupied_bits |= bit
                board.player_bits[Player(v)] |= bit
    board.rebuild_frontiers()
    return board


def think(payload: Dict[str, Any], iterations: int = 5000, leaf_batch: int = 8) -> Dict[str, Any]:
    state = payload["gameState"]
    budget = int(payload.get("timeBudgetMs", 1))
    board = rebuild_board(state)
    player = Player[state.get("current_player", "RED")]
    legal: List[Move] = [Move(m["piece_id"], m["orientation"], m["anchor_row"], m["anchor_col"])
                         for m in payload.get("legalMoves", [])]
    # the service's agent batches its leaf rollouts (8 warps per tree): the 5 000-simulation cap of the reference's
    # service (engine-service/app.py:59) is reached in ~50 ms instead of ~400 ms
    agent = FastMCTSAgent(iterations=iterations, time_limit=max(budget, 1) / 1000.0, leaf_batch=leaf_batch)
    result = agent.think(board, player, legal, budget)
    move = result["move"]
    return {
        "move": None if move is None else {"piece_id": move.piece_id, "orientation": move.orientation,
                                           "anchor_row": move.anchor_row, "anchor_col": move.anchor_col},
        "stats": result.get("stats") or {},
    }

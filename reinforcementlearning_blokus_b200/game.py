"""``BlokusGame`` with the reference's interface (``engine/game.py``), rules evaluated on the GPU.

Followed: ``engine/game.py:26-38`` (GameResult), ``:48-55`` (construction), ``:57-180`` (make_move:
legality -> place -> history -> game-over check), ``:182-214`` (game over iff nobody can move),
``:216-260`` (result / winners), ``:284-349`` (score with corner and centre bonuses), ``:351-428``
(getters).  Passing is done by the caller exactly as with the reference:
``game.board._update_current_player(); game._check_game_over()`` (arena_runner.py:662-666).

Not reproduced: the optional per-move telemetry (``enable_telemetry``; UI analytics, SURVEY.md Q11 and
section 2 "OUT OF SCOPE").  The constructor accepts the arguments and ignores them; history entries carry
the same keys minus ``metrics``/``telemetry``.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

from . import single
from .board import Board, Player, Position
from .move_generator import LegalMoveGenerator, Move, get_shared_generator


@dataclass
class GameResult:
    scores: Dict[int, int]      # Player.value -> BlokusGame score
    winner_ids: List[int]       # every Player.value at the maximum
    is_tie: bool


class BlokusGame:
    def __init__(self, enable_telemetry: bool = False, telemetry_fast_mode: bool = True):
        self.board = Board()
        self.move_generator: LegalMoveGenerator = get_shared_generator()
        self.game_history: List[dict] = []
        self.winner: Optional[Player] = None
        # engine/game.py:44-48: with telemetry every move records the four players' metrics before and after it
        # (telemetry.collect_all_player_metrics: four device launches per board) and their deltas
        self.enable_telemetry = bool(enable_telemetry)
        self.telemetry_fast_mode = telemetry_fast_mode

    # ---- moves ------------------------------------------------------------------------------------------
    def make_move(self, move: Move, player: Optional[Player] = None) -> bool:
        """False for an illegal move (board untouched), True after applying a legal one."""
        if player is None:
            player = self.board.current_player
        # one device launch: legality of the move and, after it, who can still move (engine/game.py:79-90, 182-214)
        move_id = move.to_id()            # None for bad piece / orientation indices and off-board anchors
        if move_id is None or move.piece_id in self.board.player_pieces_used[player]:
            return False
        before = None
        if self.enable_telemetry:
            from . import telemetry

            before = telemetry.collect_all_player_metrics(self.board, self.move_generator, self.telemetry_fast_mode)
        ok, alive = single.step_check(self.board.to_packed(), player.value - 1, move_id)
        if not ok:
            return False
        cells = self.move_generator.piece_position_cache[move.piece_id][move.orientation]
        positions = [Position(move.anchor_row + r, move.anchor_col + c) for r, c in cells]
        if not self.board.place_piece(positions, player, move.piece_id, validate=False):
            return False
        idx = len(self.game_history)
        self.game_history.append({
            "turn_number": idx + 1, "move_index": idx, "round_index": idx // 4, "position_in_round": idx % 4,
            "seat_index": player.value - 1, "player_to_move": player.name,
            "action": {"piece_id": move.piece_id, "orientation": move.orientation,
                       "anchor_row": move.anchor_row, "anchor_col": move.anchor_col},
            "board_state": self.board.grid.tolist(),
        })
        if before:
            after = telemetry.collect_all_player_metrics(self.board, self.move_generator, self.telemetry_fast_mode)
            payload = telemetry.compute_move_telemetry_delta(
                ply=idx + 1, mover_id=player.name,
                move_id=f"{move.piece_id}-{move.orientation}-{move.anchor_row}-{move.anchor_col}", before=before, after=after)
            if payload:
                payload["before"] = [{"playerId": pid, "metrics": m} for pid, m in before.items()]
                payload["after"] = [{"playerId": pid, "metrics": m} for pid, m in after.items()]
                self.game_history[-1]["telemetry"] = payload
        if alive == 0:
            self._set_game_over()
        return True

    def _check_game_over(self) -> None:
        """Game over iff no player has a legal move: one has-moves launch for all four players."""
        if single.has_moves_mask(self.board.to_packed()) == 0:
            self._set_game_over()

    def _set_game_over(self) -> None:
        self.board.game_over = True
        result = self.get_game_result()
        self.winner = None if result.is_tie else Player(result.winner_ids[0])

    # ---- results ------------------------------------------------------------------------------------------
    def _scores(self) -> List[int]:
        return single.scores(self.board.to_packed())[0]

    def get_game_result(self) -> GameResult:
        sc = self._scores()
        scores = {p.value: int(sc[p.value - 1]) for p in Player}
        top = max(scores.values())
        winners = [pid for pid, s in scores.items() if s == top]
        return GameResult(scores=scores, winner_ids=winners, is_tie=len(winners) > 1)

    def get_winner(self) -> Optional[Player]:
        if not self.board.game_over:
            return None
        result = self.get_game_result()
        return None if result.is_tie else Player(result.winner_ids[0])

    def get_score(self, player: Player) -> int:
        """Board score + 5 per board corner owned + 2 per owned cell in rows/cols 8..11 (engine/game.py:284-349)."""
        return int(self._scores()[player.value - 1])

    # ---- getters ------------------------------------------------------------------------------------------
    def get_legal_moves(self, player: Optional[Player] = None) -> List[Move]:
        if player is None:
            player = self.board.current_player
        return self.move_generator.get_legal_moves(self.board, player)

    def get_game_state(self) -> Dict:
        sc = self._scores()
        return {
            "board": self.board, "current_player": self.board.current_player, "move_count": self.board.move_count,
            "game_over": self.board.game_over, "winner": self.winner,
            "scores": {p.name: int(sc[p.value - 1]) for p in Player},
            "legal_moves": len(self.get_legal_moves()), "game_history_length": len(self.game_history),
        }

    def reset_game(self) -> None:
        self.board = Board()
        self.game_history = []
        self.winner = None

    def get_board_copy(self) -> Board:
        return self.board.copy()

    def is_game_over(self) -> bool:
        return self.board.game_over

    def get_current_player(self) -> Player:
        return self.board.current_player

    def get_move_count(self) -> int:
        return self.board.move_count

    @property
    def move_count(self) -> int:
        return self.board.move_count

    def get_player_pieces_used(self, player: Player) -> int:
        return len(self.board.player_pieces_used[player])

    def get_player_pieces_remaining(self, player: Player) -> int:
        return 21 - self.get_player_pieces_used(player)

    def can_player_move(self, player: Player) -> bool:
        return self.move_generator.has_legal_moves(self.board, player)

    def get_game_summary(self) -> str:
        sc = self._scores()
        lines = [f"Current Player: {self.board.current_player.name}", f"Move Count: {self.board.move_count}",
                 f"Game Over: {self.board.game_over}"]
        if self.winner:
            lines.append(f"Winner: {self.winner.name}")
        lines.append("\nScores:")
        for p in Player:
            lines.append(f"  {p.name}: {int(sc[p.value - 1])} points ({self.get_player_pieces_used(p)}/21 pieces used)")
        lines.append(f"\nLegal moves for current player: {len(self.get_legal_moves())}")
        return "\n".join(lines)

"""B200-native Blokus simulation core: a drop-in for the hot path of TGALLOWAY1/ReinforcementLearning_Blokus.

Object API (same names and semantics as the reference's ``engine`` / ``agents`` / ``mcts`` packages):
    Board, Player, Position, Move, LegalMoveGenerator, get_shared_generator, BlokusGame, GameResult,
    RandomAgent, HeuristicAgent, FastMCTSAgent, GameplayFastMCTSAgent, MCTSAgent, ZobristHash, TranspositionTable
Batched device API (the throughput path): ``device`` (tensors in/out), ``batched.BatchedBlokus``, ``mcts.DeviceUCT``,
``dist`` (game sharding and root-parallel search across GPUs).

Importing this package needs neither a GPU nor the compiled library; the first compute call loads
``libblokus_b200.so`` and raises if it or a CUDA device is missing (there is no CPU fallback).
"""

from .pieces import BOARD_SIZE, NUM_ORIENTATIONS, NUM_PIECES, ORIENTATIONS, PASS_ID, decode_move, encode_move  # noqa: F401
from .state import STATE_BYTES, STATE_DTYPE, empty_states  # noqa: F401

_LAZY = {
    "Board": "board", "Player": "board", "Position": "board",
    "Move": "move_generator", "LegalMoveGenerator": "move_generator", "get_shared_generator": "move_generator",
    "BlokusGame": "game", "GameResult": "game",
    "RandomAgent": "agents", "HeuristicAgent": "agents", "FastMCTSAgent": "agents", "GameplayFastMCTSAgent": "agents",
    "MCTSAgent": "mcts", "ZobristHash": "mcts", "TranspositionTable": "mcts", "DeviceUCT": "mcts",
    "BatchedBlokus": "batched", "LearnedWinProbabilityEvaluator": "winprob",
}


_SUBMODULES = ("agents", "arena", "batched", "board", "device", "dist", "game", "integration", "mcts", "move_generator",
               "mobility", "service", "single", "telemetry", "winprob")


def __getattr__(name):
    import importlib

    if name in _SUBMODULES:
        return importlib.import_module(f".{name}", __name__)
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError(name)
    return getattr(importlib.import_module(f".{mod}", __name__), name)


__all__ = sorted(_LAZY) + ["BOARD_SIZE", "NUM_ORIENTATIONS", "NUM_PIECES", "ORIENTATIONS", "PASS_ID", "decode_move",
                           "encode_move", "STATE_BYTES", "STATE_DTYPE", "empty_states"]
__version__ = "0.1.0"

"""Where the host time of FastMCTSAgent.think goes (cProfile over 200 calls on a 237-move position)."""
import cProfile
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning_blokus_b200 import Board, FastMCTSAgent, get_shared_generator  # noqa: E402
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.state import STATE_DTYPE  # noqa: E402

mid = dev.playout(64, seed=0xB10C05, max_turns=24, want_final=True)["final_states"][17:18].contiguous()
board = Board.from_packed(np.frombuffer(mid.cpu().numpy().tobytes(), dtype=STATE_DTYPE)[0])
player = board.current_player
legal = get_shared_generator().get_legal_moves(board, player)
agent = FastMCTSAgent(iterations=1000, time_limit=30.0, seed=1)
for _ in range(5):
    agent.think(board, player, legal, 30000)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    agent.think(board, player, legal, 30000)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

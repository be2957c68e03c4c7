"""Latency of the object-API agents on one board (BASELINE config 4: 1,000 simulations per move)."""
import sys, time, json
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import BlokusGame, FastMCTSAgent, MCTSAgent, RandomAgent, get_shared_generator

game = BlokusGame()
rnd = RandomAgent(seed=1)
for _ in range(16):                                  # a mid-game root
    p = game.get_current_player()
    game.make_move(rnd.select_action(game.board, p, game.get_legal_moves(p)), p)
board, player = game.board, game.get_current_player()
gen = get_shared_generator()
legal = gen.get_legal_moves(board, player)
out = {"root_moves": len(legal)}
t = time.perf_counter()
for _ in range(50): gen.get_legal_moves(board, player)
out["get_legal_moves_us"] = (time.perf_counter() - t) / 50 * 1e6
for name, agent in (("mcts_uniform_1tree", MCTSAgent(iterations=1000, seed=0, rollout_agent=RandomAgent(0))),
                    ("mcts_heuristic_1tree", MCTSAgent(iterations=1000, seed=0)),
                    ("mcts_uniform_64trees", MCTSAgent(iterations=1000, seed=0, rollout_agent=RandomAgent(0), num_trees=64,
                                                       use_transposition_table=False)),
                    ("fast_mcts_1000", FastMCTSAgent(iterations=1000, time_limit=60.0, seed=0)),
                    # GPU-batched leaf rollouts: the same 1,000 simulations, 8 per iteration (a CTA of 8 warps per tree)
                    ("fast_mcts_1000_leaf_batch8", FastMCTSAgent(iterations=1000, time_limit=60.0, seed=0, leaf_batch=8)),
                    ("fast_mcts_1000_leaf_batch16", FastMCTSAgent(iterations=1000, time_limit=60.0, seed=0, leaf_batch=16)),
                    ("mcts_heuristic_125x8_leaf_batch8", MCTSAgent(iterations=125, seed=0, leaf_batch=8,
                                                                  use_transposition_table=False)),
                    ("mcts_uniform_1000x8_leaf_batch8", MCTSAgent(iterations=1000, seed=0, rollout_agent=RandomAgent(0),
                                                                 leaf_batch=8, use_transposition_table=False))):
    agent.select_action(board, player, legal)         # warm-up (allocations, lazy loads)
    t = time.perf_counter()
    mv = agent.select_action(board, player, legal)
    dt = time.perf_counter() - t
    sims = {"mcts_heuristic_125x8_leaf_batch8": 1000, "mcts_uniform_1000x8_leaf_batch8": 8000}.get(
        name, 1000 * getattr(agent, "num_trees", 1))
    out[name] = {"ms_per_move": dt * 1e3, "sims_per_sec": sims / dt, "move": str(mv)}
print(json.dumps(out, indent=1))

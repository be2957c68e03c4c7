"""Latency of ONE move searched by the wide (round-synchronous) kernel vs one warp / one CTA (profiles/r02_wide_latency.json)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.mcts import DeviceUCT, WideUCT  # noqa: E402
from reinforcementlearning_blokus_b200.state import empty_states  # noqa: E402


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return round(min(ts), 3), round(float(np.median(ts)), 3)


def main():
    out = {}
    mid = dev.playout(64, seed=0xB10C05, max_turns=24, want_final=True)["final_states"][17:18].contiguous()
    roots = {"empty": dev.to_device(empty_states(1)), "turn24": mid}
    flat = dict(max_depth=1, ucb_variant=1, reward_mode=1)
    for rname, root in roots.items():
        n_legal = int(dev.movegen_count(root)[0])
        row = {"n_legal": n_legal}
        for sims in (1000, 5000):
            for width in (128, 256, 512, 1024, 2048):
                eng = WideUCT(sims, width=width, seed=1, **flat)
                row[f"flat_{sims}_w{width}_ms(min,med)"] = timed(lambda: eng.search(root, sims))
            for width in (128, 256, 512):
                eng = WideUCT(sims, width=width, seed=1)
                row[f"uct_{sims}_w{width}_ms(min,med)"] = timed(lambda: eng.search(root, sims))
        for width in (64, 128, 256, 512):
            eng = WideUCT(1000, width=width, seed=1, rollout_policy=1)
            row[f"uct_heuristic_1000_w{width}_ms(min,med)"] = timed(lambda: eng.search(root, 1000), reps=3)
        for trees in (2, 4, 8):
            per = 5000 // trees
            eng = WideUCT(per, width=max(64, min(1024, per // 4)), n_trees=trees, seed=1, **flat)
            row[f"flat_5000_in_{trees}_trees_ms(min,med)"] = timed(lambda: eng.search(root, per))
        eng = DeviceUCT(1000, n_trees=1, seed=1, **flat)
        row["flat_1000_one_warp_ms(min,med)"] = timed(lambda: eng.search(root, 1000), reps=2)
        eng = DeviceUCT(63, n_trees=1, seed=1, leaf_batch=16, **flat)
        row["flat_1008_leaf_batch16_ms(min,med)"] = timed(lambda: eng.search(root, 63), reps=3)
        eng = DeviceUCT(1000, n_trees=1, seed=1, rollout_policy=1)
        row["uct_heuristic_1000_one_warp_ms(min,med)"] = timed(lambda: eng.search(root, 1000), reps=2)
        out[rname] = row
    # the agents end to end (object API: packing the board, the search, reading the statistics back)
    from reinforcementlearning_blokus_b200 import Board, FastMCTSAgent, MCTSAgent, get_shared_generator
    from reinforcementlearning_blokus_b200.state import STATE_DTYPE
    board = Board.from_packed(mid.cpu().numpy().view(STATE_DTYPE)[0, 0] if False else np.frombuffer(mid.cpu().numpy().tobytes(), dtype=STATE_DTYPE)[0])
    player = board.current_player
    legal = get_shared_generator().get_legal_moves(board, player)
    ag = {}
    for name, agent in (("FastMCTSAgent_think_1000", FastMCTSAgent(iterations=1000, time_limit=30.0, seed=1)),
                        ("FastMCTSAgent_think_5000", FastMCTSAgent(iterations=5000, time_limit=30.0, seed=1)),
                        ("MCTSAgent_1000_heuristic_wide256", MCTSAgent(iterations=1000, seed=1, search_width=256, use_transposition_table=False)),
                        ("MCTSAgent_1000_heuristic_sequential", MCTSAgent(iterations=1000, seed=1))):
        fn = (lambda a=agent: a.think(board, player, legal, 30000)) if hasattr(agent, "think") else (lambda a=agent: a.select_action(board, player, legal))
        ag[name + "_ms(min,med)"] = timed(fn, reps=3)
    out["agents_turn24"] = ag
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/r02_wide_latency.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

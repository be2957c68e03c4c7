"""Throughput of MCTS self-play (every move of every game searched): games/s and simulations/s on one GPU."""
import sys
import time
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import batched, device as dev

for n, iters, pol in ((2048, 64, dev.POLICY_UNIFORM), (1024, 256, dev.POLICY_UNIFORM), (512, 64, dev.POLICY_HEURISTIC)):
    batched.selfplay_mcts(64, 8, seed=1, max_turns=4)                    # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = batched.selfplay_mcts(n, iters, seed=7, rollout_policy=pol)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    placed = int((r["trace"].cpu().numpy() != 0xFFFF).sum())
    print({"games": n, "iterations_per_move": iters, "rollout_policy": pol, "seconds": round(dt, 2), "games_per_sec": round(n / dt, 1),
           "searched_moves_per_sec": round(placed / dt, 1), "sims_per_sec": round(placed * iters / dt),
           "mean_score": float(r["game_scores"].float().mean())})

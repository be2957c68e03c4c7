"""Latency of one UCT move with HeuristicAgent rollouts on the wide search (tuning aid; BLOKUS_B200_LIB selects the build)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import WideUCT
mid = dev.playout(64, seed=0xB10C05, max_turns=24, want_final=True)["final_states"][17:18].contiguous()
out = {"lib": os.environ.get("BLOKUS_B200_LIB", "default")}
for width in (64, 256, 512):
    eng = WideUCT(1000, width=width, seed=1, rollout_policy=1)
    eng.search(mid, 1000); eng.search(mid, 1000); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); eng.search(mid, 1000); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    out[f"uct_heuristic_1000_w{width}_ms"] = round(min(ts), 3)
print(json.dumps(out))

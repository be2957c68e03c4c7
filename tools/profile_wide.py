"""One wide search per configuration (for ncu launch lists / full captures of k_wide_plan and k_wide_work)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.mcts import WideUCT  # noqa: E402
from reinforcementlearning_blokus_b200.state import empty_states  # noqa: E402

sims = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
width = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
mode = sys.argv[3] if len(sys.argv) > 3 else "flat"
root = dev.playout(64, seed=0xB10C05, max_turns=24, want_final=True)["final_states"][17:18].contiguous()
kw = dict(max_depth=1, ucb_variant=1, reward_mode=1) if mode == "flat" else (dict(rollout_policy=1) if mode == "heur" else {})
eng = WideUCT(sims, width=width, seed=1, **kw)
for _ in range(2):
    eng.search(root, sims)
torch.cuda.synchronize()
print("ok", int(eng.done[0]))

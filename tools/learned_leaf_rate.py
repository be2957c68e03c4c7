"""UCT with learned leaf evaluation (reward_mode 3): simulations per second, many trees and one tree."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import DeviceUCT
from reinforcementlearning_blokus_b200.state import empty_states

w = torch.from_numpy(np.random.default_rng(0).normal(0, 0.05, 35))
roots = dev.to_device(empty_states(1))
for trees, iters in ((2368, 256), (9472, 256), (1, 1000)):
    eng = DeviceUCT(iters, n_trees=trees, seed=42, reward_mode=3, leaf_model=w)
    eng.search(roots, iters); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.search(roots, iters); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print({"trees": trees, "iterations": iters, "ms": round(ms, 2), "sims_per_sec": round(trees * iters / ms * 1e3)})
    del eng

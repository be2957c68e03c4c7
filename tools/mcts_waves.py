"""UCT throughput against the number of root-parallel trees (1, 2, 4 waves of 148 SMs x 24 resident warps)."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import DeviceUCT
from reinforcementlearning_blokus_b200.state import empty_states

roots = dev.to_device(empty_states(1))
for pol, per in ((dev.POLICY_UNIFORM, 3552), (dev.POLICY_HEURISTIC, 2072)):
    for waves in (1, 2, 4):
        trees = per * waves
        eng = DeviceUCT(256, n_trees=trees, seed=42, rollout_policy=pol)
        eng.search(roots, 256)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.search(roots, 256); b.record(); torch.cuda.synchronize()
        print({"policy": pol, "trees": trees, "sims_per_sec": trees * 256 / a.elapsed_time(b) * 1e3}, flush=True)
        del eng

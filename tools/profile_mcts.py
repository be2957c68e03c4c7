"""ncu driver: one root-parallel UCT search per rollout policy (bench settings: 3552 trees x 256 iterations, 50 plies)."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import DeviceUCT
from reinforcementlearning_blokus_b200.state import empty_states

roots = dev.to_device(empty_states(1))
for pol, trees in ((dev.POLICY_UNIFORM, 3552), (dev.POLICY_HEURISTIC, 1184)):
    eng = DeviceUCT(256, n_trees=trees, seed=42, rollout_policy=pol)
    eng.search(roots, 256)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.search(roots, 256); b.record(); torch.cuda.synchronize()
    print("policy", pol, "trees", trees, "sims/s", trees * 256 / a.elapsed_time(b) * 1e3)

"""ncu driver: one short launch of k_mcts<1,0,0> (UCT with HeuristicAgent rollouts), whose counter passes came back as nan
inside the long all-kernels capture."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import DeviceUCT
from reinforcementlearning_blokus_b200.state import empty_states

roots = dev.to_device(empty_states(1))
eng = DeviceUCT(8, n_trees=2072, seed=42, rollout_policy=dev.POLICY_HEURISTIC)
eng.search(roots, 8)
eng.search(roots, 8)
torch.cuda.synchronize()
print("ok")

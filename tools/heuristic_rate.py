"""Heuristic-policy playouts/s (16,384 games per launch, 5 launches) and UCT sims/s with heuristic rollouts."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import DeviceUCT
from reinforcementlearning_blokus_b200.state import empty_states

out = {}
for _ in range(3):
    dev.playout(16384, seed=42, policy=1, out=out, want_board_scores=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    dev.playout(16384, seed=42, policy=1, out=out, want_board_scores=False)
b.record(); torch.cuda.synchronize()
print({"heuristic_playouts_per_sec": 5 * 16384 / a.elapsed_time(b) * 1e3})
roots = dev.to_device(empty_states(1))
eng = DeviceUCT(256, n_trees=4144, seed=42, rollout_policy=dev.POLICY_HEURISTIC)
eng.search(roots, 256); torch.cuda.synchronize()
a.record(); eng.search(roots, 256); b.record(); torch.cuda.synchronize()
print({"heuristic_rollout_sims_per_sec": 4144 * 256 / a.elapsed_time(b) * 1e3})

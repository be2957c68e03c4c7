"""Small driver for ncu captures of the non-headline kernels (movegen count / emit / bitmap, UCT search)."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import DeviceUCT
from reinforcementlearning_blokus_b200.state import empty_states

n = 65536
parts = [dev.playout(n // 4, seed=7, game_offset=k * 1_000_000, max_turns=k, want_final=True)["final_states"] for k in (8, 16, 24, 32, 40, 48)]
corpus = torch.cat(parts).contiguous()
for _ in range(3):
    dev.movegen_count(corpus)
    dev.movegen(corpus)
    dev.movegen_bitmap(corpus)
eng = DeviceUCT(128, n_trees=1184, seed=42)
roots = dev.to_device(empty_states(1))
for _ in range(2):
    eng.search(roots, 128)
torch.cuda.synchronize()
print("ok")

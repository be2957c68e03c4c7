"""ncu driver: one uniform and one heuristic playout batch (65,536 / 16,384 games)."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
out = {}
for _ in range(3):
    dev.playout(65536, seed=42, out=out, want_board_scores=False)
out2 = {}
for _ in range(2):
    dev.playout(16384, seed=42, policy=1, out=out2, want_board_scores=False)
torch.cuda.synchronize()
print("ok")

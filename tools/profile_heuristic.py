import sys, torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
out = {}
for _ in range(3):
    dev.playout(16384, seed=42, policy=1, out=out, want_board_scores=False)
torch.cuda.synchronize(); print("ok")

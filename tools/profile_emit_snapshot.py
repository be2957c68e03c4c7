"""ncu driver: one launch each of k_movegen_emit (CSR), k_snapshot_features and k_replay_snapshots on a mid-game corpus."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev

n = 1 << 18
parts = [dev.playout(n // 8, seed=7, game_offset=k * 1_000_000, max_turns=k, want_final=True)["final_states"].clone()
         for k in (6, 12, 18, 24, 30, 36, 42, 48)]
corpus = torch.cat(parts).contiguous()
tr = dev.playout(65536, seed=42, want_trace=True)["trace"].clone()
dev.movegen(corpus)
dev.snapshot_features(corpus)
dev.replay_snapshots(tr, (8, 16, 24, 32, 40, 48, 56, 64))
torch.cuda.synchronize()
print("ok")

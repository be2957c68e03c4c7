"""Snapshot feature rows per second on a 2^18-state mid-game corpus, and recorded-game replay with 8 checkpoints."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev

n = 1 << 18
parts = [dev.playout(n // 8, seed=7, game_offset=k * 1_000_000, max_turns=k, want_final=True)["final_states"].clone()
         for k in (6, 12, 18, 24, 30, 36, 42, 48)]
corpus = torch.cat(parts).contiguous()
tr = dev.playout(65536, seed=42, want_trace=True)["trace"].clone()


def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print({"snapshot_boards_per_sec": n / timeit(lambda: dev.snapshot_features(corpus)) * 1e3,
       "snapshot_replay_games_per_sec": 65536 / timeit(lambda: dev.replay_snapshots(tr, (8, 16, 24, 32, 40, 48, 56, 64))) * 1e3})

"""CSR / bitmap / has_moves / heuristic rates in one go (tuning aid for the row-sink generators)."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev

n = 1 << 20
parts = [dev.playout(n // 8, seed=7, game_offset=k * 1_000_000, max_turns=k, want_final=True)["final_states"].clone()
         for k in (6, 12, 18, 24, 30, 36, 42, 48)]
corpus = torch.cat(parts).contiguous()


def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {}
print({"csr_G_per_s": round(n / timeit(lambda: dev.movegen(corpus)) / 1e6, 4),
       "bitmap_G_per_s": round(n / timeit(lambda: dev.movegen_bitmap(corpus)) / 1e6, 4),
       "has_moves_G_per_s": round(n / timeit(lambda: dev.has_moves(corpus)) / 1e6, 4),
       "heuristic_playouts_M_per_s": round(16384 / timeit(lambda: dev.playout(16384, seed=42, policy=1, out=out, want_board_scores=False), reps=3) / 1e3, 4),
       "checksum": int(dev.movegen(corpus)[2].long().sum())})

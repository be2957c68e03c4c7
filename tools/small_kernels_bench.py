"""Rates of the secondary kernels on a 2^20-state mid-game corpus (bitmap, has_moves, step, scores, zobrist, is_move_legal)."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import ZobristHash

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


ztab = ZobristHash(seed=1).device_table()
counts, offsets, ids = dev.movegen(corpus)
first = ids.view(torch.int16)[offsets[:-1].long().clamp(max=ids.numel() - 1)].contiguous().view(torch.uint16)       # first legal move of every state
work = corpus.clone()
out = {
    "bitmap_G_per_s": n / timeit(lambda: dev.movegen_bitmap(corpus)) / 1e6,
    "has_moves_G_per_s": n / timeit(lambda: dev.has_moves(corpus)) / 1e6,
    "scores_G_per_s": n / timeit(lambda: dev.scores(corpus)) / 1e6,
    "zobrist_G_per_s": n / timeit(lambda: dev.zobrist_hash(corpus, ztab)) / 1e6,
    "is_move_legal_G_per_s": n / timeit(lambda: dev.is_move_legal(corpus, first)) / 1e6,
    "step_with_alive_G_per_s": n / timeit(lambda: (work.copy_(corpus), dev.step(work, first, want_alive=True))) / 1e6,
}
print({k: round(v, 3) for k, v in out.items()})

"""Quick device timing of the hot kernels (development aid; bench.py is the contract)."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
out = {}
ms = timeit(lambda: dev.playout(n, seed=42, out=out, want_board_scores=False))
print(f"playout n={n}: {ms:.2f} ms -> {n/ms*1e3/1e6:.3f} M playouts/s")
r = dev.playout(n, seed=42, want_trace=True, want_final=True)
placed = (r["trace"].cpu().numpy() != 0xFFFF).sum(axis=1)
print("mean placements/game", placed.mean(), "max", placed.max())
# mid-game corpus: stop after k turns
states = []
for k in (8, 16, 24, 32, 40, 48):
    states.append(dev.playout(n // 4, seed=7, game_offset=k * 100000, max_turns=k, want_final=True)["final_states"])
corpus = torch.cat(states).contiguous()
m = corpus.shape[0]
ms = timeit(lambda: dev.movegen_count(corpus))
print(f"movegen_count n={m}: {ms:.3f} ms -> {m/ms*1e3/1e6:.1f} M gens/s")
ms = timeit(lambda: dev.movegen(corpus))
print(f"movegen CSR n={m}: {ms:.3f} ms -> {m/ms*1e3/1e6:.1f} M gens/s; mean legal {dev.movegen_count(corpus).float().mean().item():.1f}")
ms = timeit(lambda: dev.movegen_bitmap(corpus))
print(f"movegen bitmap n={m}: {ms:.3f} ms -> {m/ms*1e3/1e6:.1f} M gens/s")

"""Times the playout kernel of whatever library BLOKUS_B200_LIB points at (tuning aid)."""
import sys, os
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
n = 65536
out = {}
for _ in range(3): dev.playout(n, seed=42, out=out, want_board_scores=False)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(10): dev.playout(n, seed=42, game_offset=k * n, out=out, want_board_scores=False)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
print(os.environ.get("BLOKUS_B200_LIB", "default"), f"{best:.3f} ms/step -> {n/best*1e3/1e6:.2f} M playouts/s",
      "checksum", int(out["game_scores"].sum()))

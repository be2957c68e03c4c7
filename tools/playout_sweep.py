#!/usr/bin/env python
"""Replay parity at scale: N whole games played by the CUDA kernel are replayed move for move by the oracle
(oracle/blokus_oracle.c: the reference's engine + the same Philox index stream) and must agree on every move id of
the trace and on the four final BlokusGame scores.  Prints one JSON line."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import orc  # noqa: E402
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402


def sweep(n: int, seed: int, policy: int, first_game: int = 0) -> dict:
    t0 = time.perf_counter()
    r = dev.playout(n, seed=seed, game_offset=first_game, policy=policy, want_trace=True)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    tr, gs = r["trace"].cpu().numpy(), r["game_scores"].cpu().numpy()
    t0 = time.perf_counter()
    bad, first = orc.check_playouts(seed, first_game, tr, gs, policy=policy)
    return {"policy": "heuristic" if policy else "uniform", "games": n, "seed": seed, "mismatches": int(bad),
            "first_bad": int(first), "placements": int((tr != 0xFFFF).sum()), "mean_score": float(gs.mean()),
            "gpu_seconds": round(t_gpu, 4), "oracle_seconds": round(time.perf_counter() - t0, 1)}


if __name__ == "__main__":
    n_uniform = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n_heur = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    print(json.dumps([sweep(n_uniform, 42, 0, first_game=10_000_000), sweep(n_heur, 42, 1, first_game=20_000_000)]))

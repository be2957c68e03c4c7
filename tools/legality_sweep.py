#!/usr/bin/env python
"""BASELINE.json configs[2]: legality bit-exactness sweep over N synthetic states (default 10^6).

States are produced on the device: game g of group k is uniform-random self-play from the empty board stopped
after k turns (k = 0..60, so every game stage is covered, terminal states included).  For every state the
CUDA generator's ascending move list for the side to move is compared, byte for byte, with the oracle's
(oracle/blokus_oracle.c: the reference's default frontier+bitboard generator; every 1000th state also against
its naive generator).  Prints one JSON line.
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import orc  # noqa: E402
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.state import STATE_DTYPE  # noqa: E402


def sweep(n_total: int, seed: int = 0xB10C05) -> dict:
    from reinforcementlearning_blokus_b200.state import empty_states

    groups = 61
    per = -(-n_total // groups)
    parts = [dev.to_device(empty_states(per))]                       # k = 0: the empty board
    for k in range(1, groups):
        r = dev.playout(per, seed=seed, game_offset=k * per, max_turns=k, want_final=True)
        parts.append(r["final_states"].clone())
    states = torch.cat(parts)[:n_total].contiguous()
    t0 = time.perf_counter()
    counts, offsets, ids = dev.movegen(states)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    st = states.cpu().numpy().view(STATE_DTYPE).reshape(-1)
    off = offsets.cpu().numpy().view(np.uint32)
    mv = ids.cpu().numpy()
    t0 = time.perf_counter()
    bad, first = orc.check_moves_batch(st, off, mv)
    t_cpu = time.perf_counter() - t0
    sub = np.arange(0, len(st), 1000)
    sub_off = np.concatenate([[0], np.cumsum(counts.cpu().numpy()[sub])]).astype(np.uint32)
    sub_ids = np.concatenate([mv[off[i]:off[i + 1]] for i in sub]) if len(sub) else mv[:0]
    bad_naive, _ = orc.check_moves_batch(st[sub], sub_off, sub_ids, naive=True)
    return {"states": int(len(st)), "total_moves": int(off[-1]), "mismatches": int(bad), "first_bad": int(first),
            "naive_subsample": int(len(sub)), "naive_mismatches": int(bad_naive),
            "terminal_states": int((counts == 0).sum()), "max_moves": int(counts.max()),
            "gpu_seconds": round(t_gpu, 3), "oracle_seconds": round(t_cpu, 1)}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # chunks of n states, each with its own seed
    if chunks == 1:
        print(json.dumps(sweep(n)))
    else:
        tot = None
        for c in range(chunks):
            r = sweep(n, seed=0xB10C05 + 7919 * c)
            torch.cuda.empty_cache()
            if tot is None:
                tot = dict(r, chunks=1)
            else:
                for k in ("states", "total_moves", "mismatches", "naive_subsample", "naive_mismatches", "terminal_states"):
                    tot[k] += r[k]
                tot["max_moves"] = max(tot["max_moves"], r["max_moves"])
                tot["gpu_seconds"] = round(tot["gpu_seconds"] + r["gpu_seconds"], 3)
                tot["oracle_seconds"] = round(tot["oracle_seconds"] + r["oracle_seconds"], 1)
                tot["chunks"] += 1
                if tot["first_bad"] < 0 and r["first_bad"] >= 0:
                    tot["first_bad"] = r["first_bad"]
        print(json.dumps(tot))

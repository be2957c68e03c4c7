"""Latency of ONE move searched by the wide (round-synchronous) kernel vs one warp / one CTA (profiles/r02_wide_latency.json)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.mcts import DeviceUCT, WideUCT  # noqa: E402
from reinforcementlearning_blokus_b200.state import empty_states  # noqa: E402


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return round(min(ts), 3), round(float(np.median(ts)), 3)


def main():
    out = {}
    mid = dev.playout(64, seed=0xB10C05, max_turns=24, want_final=True)["final_states"][17:18].contiguous()
    roots = {"empty": dev.to_device(empty_states(1)), "turn24": mid}
    flat = dict(max_depth=1, ucb_variant=1, reward_mode=1)
    for rname, root in roots.items():
        n_legal = int(dev.movegen_count(root)[0])
        row = {"n_legal": n_legal}
        for sims in (1000, 5000):
            for width in (128, 256, 512, 1024, 2048):
                eng = WideUCT(sims, width=width, seed=1, **flat)
                row[f"flat_{sims}_w{width}_ms(min,med)"] = timed(lambda: eng.search(root, sims))
            for width in (128, 256, 512):
                eng = WideUCT(sims, width=width, seed=1)
                row[f"uct_{sims}_w{width}_ms(min,med)"] = timed(lambda: eng.search(root, sims))
        for width in (64, 128, 256, 512):
            eng = WideUCT(1000, width=width, seed=1, rollout_policy=1)
            row[f"uct_heuristic_1000_w{width}_ms(min,med)"] = timed(lambda: eng.search(root, 1000), reps=3)
        eng = DeviceUCT(1000, n_trees=1, seed=1, **flat)
        row["flat_1000_one_warp_ms(min,med)"] = timed(lambda: eng.search(root, 1000), reps=2)
        eng = DeviceUCT(63, n_trees=1, seed=1, leaf_batch=16, **flat)
        row["flat_1008_leaf_batch16_ms(min,med)"] = timed(lambda: eng.search(root, 63), reps=3)
        eng = DeviceUCT(1000, n_trees=1, seed=1, rollout_policy=1)
        row["uct_heuristic_1000_one_warp_ms(min,med)"] = timed(lambda: eng.search(root, 1000), reps=2)
        out[rname] = row
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/r02_wide_latency.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

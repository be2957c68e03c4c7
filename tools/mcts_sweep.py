#!/usr/bin/env python
"""UCT parity at scale: many roots x trees searched by the CUDA kernel and, sequentially, by the CPU oracle (the same
algorithm restated from mcts/mcts_agent.py with the same canonical untried order and Philox rollout stream); root
visit counts, reward sums and move ids must be identical for every tree.  Prints one JSON line.

  python tools/mcts_sweep.py [n_roots] [iterations]
"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import orc  # noqa: E402
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.mcts import DeviceUCT  # noqa: E402


def main(n_roots: int, iters: int) -> None:
    # roots: random self-play stopped after 0..47 turns
    parts = [dev.playout(n_roots // 8, seed=11, game_offset=k * 1000, max_turns=k, want_final=True)["final_states"].clone()
             for k in (0, 4, 9, 15, 22, 30, 39, 47)]
    roots = torch.cat(parts).contiguous()
    alive = (dev.movegen_count(roots) > 0).cpu()
    roots = roots[alive.to(roots.device)].contiguous()
    raw = roots.cpu().numpy().view(orc.STATE_DTYPE).reshape(-1)
    out = []
    for name, kw in (("uct_uniform", dict(rollout_policy=0)), ("uct_heuristic", dict(rollout_policy=1)),
                     ("flat_root_reward", dict(max_depth=1, ucb_variant=1, reward_mode=1)),
                     ("constant_leaf", dict(reward_mode=2))):
        n = len(raw) if name != "uct_heuristic" else max(1, len(raw) // 4)
        t0 = time.perf_counter()
        eng = DeviceUCT(iters, n_trees=n, seed=77, **kw)
        res = eng.search(roots[:n].contiguous(), iters)            # tree k searches root k with stream (seed, k)
        torch.cuda.synchronize()
        t_gpu = time.perf_counter() - t0
        vis, rew, mov, cnt = (res[k].cpu().numpy() for k in ("visits", "reward", "moves", "counts"))

        def check(k):
            w = orc.mcts_search(raw[k], iters, seed=77, tree_id=k, **kw)
            m = w["n_root"]
            return bool(m == cnt[k] and np.array_equal(vis[k, :m], w["visits"][:m])
                        and np.array_equal(rew[k, :m], w["reward"][:m]) and np.array_equal(mov[k, :m], w["moves"][:m]))

        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
            ok = list(pool.map(check, range(n)))
        out.append({"variant": name, "trees": n, "iterations": iters, "mismatches": int(n - sum(ok)),
                    "first_bad": next((i for i, v in enumerate(ok) if not v), -1),
                    "gpu_seconds": round(t_gpu, 3), "oracle_seconds": round(time.perf_counter() - t0, 1)})
        del eng
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 200)

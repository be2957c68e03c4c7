"""SURVEY 8d C5 on real GPUs: root-parallel search sharded over the ranks (torchrun, NCCL), one all-reduce of the root
statistics.  Every rank must hold the same merged result, visits must add up to the simulations, and the merged result
must equal ONE process searching all world_size * trees_per_rank trees itself (same global tree ids)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200 import dist as bdist
from reinforcementlearning_blokus_b200.mcts import DeviceUCT
from reinforcementlearning_blokus_b200.state import empty_states

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
root = dev.playout(1, seed=5, max_turns=12, want_final=True)["final_states"]
tpr, sims = 64, 64 * 32 * world
out = bdist.root_parallel_search(root, sims, tpr, seed=99)
v = out["visits"]
ok = int(v.sum().item()) == out["sims"]
# the same trees in one process
solo = DeviceUCT(out["sims"] // (world * tpr), n_trees=world * tpr, seed=99, tree_offset=0)
res = solo.search(root, out["sims"] // (world * tpr))
n = out["n_root"]
same = bool(torch.equal(res["visits"].sum(dim=0)[:n].to(torch.int64), v)) and bool(torch.equal(res["reward"].sum(dim=0)[:n], out["reward"]))
if world > 1:
    flag = torch.tensor([int(ok and same), out["rank_of_best"]], device="cuda")
    gathered = [torch.zeros_like(flag) for _ in range(world)]
    dist.all_gather(gathered, flag)
    agree = all(int(g[0]) == 1 and int(g[1]) == out["rank_of_best"] for g in gathered)
else:
    agree = ok and same
if rank == 0:
    print(json.dumps({"world_size": world, "trees_per_rank": tpr, "sims": out["sims"], "visits_add_up": ok,
                      "equals_single_process": same, "all_ranks_agree": bool(agree), "best_rank": out["rank_of_best"]}))
if world > 1:
    dist.destroy_process_group()

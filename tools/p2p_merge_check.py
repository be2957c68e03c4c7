"""PeerRootMerge (bk_root_merge_p2p: one kernel per rank over NVLink peer memory) against the NCCL all-reduce it replaces.
Run under torchrun on N GPUs of one box (or plainly on one GPU): many merges of random statistics with changing sizes must
equal dist.all_reduce(sum) on every rank; then both are timed.  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import dist as bdist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
peer = bdist.PeerRootMerge(capacity=2048)
g = torch.Generator(device="cuda").manual_seed(1000 + rank)
bad = 0
for it in range(200):
    n_trees = 1 + it % 5
    n = (1, 58, 237, 1003, 2048)[it % 5]
    vis = torch.randint(0, 1 << 20, (n_trees, 2048), generator=g, device="cuda", dtype=torch.int32)
    rew = torch.randint(-(1 << 40), 1 << 40, (n_trees, 2048), generator=g, device="cuda", dtype=torch.int64)
    v, w = peer.merge(vis, rew, n)
    v, w = v.clone(), w.clone()
    ref = torch.stack([vis.sum(dim=0, dtype=torch.int64), rew.sum(dim=0)])
    if world > 1:
        dist.all_reduce(ref)
    bad += int(not (torch.equal(v, ref[0, :n]) and torch.equal(w, ref[1, :n])))
# timing: the merge of a 237-move root, 8 trees per rank
vis = torch.randint(0, 1 << 20, (8, 2048), generator=g, device="cuda", dtype=torch.int32)
rew = torch.randint(0, 1 << 30, (8, 2048), generator=g, device="cuda", dtype=torch.int64)
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
t_p2p = t_nccl = 0.0
reps = 50
for _ in range(reps):
    if world > 1:
        dist.barrier()
    a.record()
    peer.merge(vis, rew, 237, check=False)
    b.record()
    bdist.merge_root_stats(vis.sum(dim=0), rew.sum(dim=0))
    c.record()
    torch.cuda.synchronize()
    t_p2p += a.elapsed_time(b)
    t_nccl += b.elapsed_time(c)
t = torch.tensor([t_p2p / reps, t_nccl / reps, float(bad), float(peer.status.item())], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world_size": world, "merges_checked": 200, "mismatching_merges_max_over_ranks": int(t[2].item()),
                      "status": int(t[3].item()), "p2p_merge_us": round(float(t[0]) * 1e3, 1),
                      "nccl_sum_convert_all_reduce_us": round(float(t[1]) * 1e3, 1)}))
peer.close()
if world > 1:
    dist.destroy_process_group()
sys.exit(1 if int(t[2].item()) or int(t[3].item()) else 0)

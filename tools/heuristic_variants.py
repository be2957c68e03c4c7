"""Heuristic-policy playouts/s and UCT-with-heuristic-rollouts sims/s of the library BLOKUS_B200_LIB points at (tuning aid).
usage: python tools/heuristic_variants.py [trees_per_sm]

A variant is built next to the product library, e.g. the self-checking build of profiles/r02p_heuristic_bounds_check.json:
  cd reinforcementlearning_blokus_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared \
      -Xcompiler -fPIC -DBK_HEUR_CHECK -o ../../build_variants/h_check.so blokus_b200.cu
  BLOKUS_B200_LIB=build_variants/h_check.so python tools/heuristic_variants.py 20"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import DeviceUCT
from reinforcementlearning_blokus_b200.state import empty_states

per_sm = int(sys.argv[1]) if len(sys.argv) > 1 else 14
out = {}
n = 32768
for _ in range(2):
    dev.playout(n, seed=42, policy=1, out=out, want_board_scores=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for k in range(4):
    dev.playout(n, seed=42, policy=1, game_offset=k * n, out=out, want_board_scores=False)
b.record(); torch.cuda.synchronize()
res = {"lib": os.environ.get("BLOKUS_B200_LIB", "default"), "heuristic_playouts_per_sec": round(4 * n / a.elapsed_time(b) * 1e3),
       "checksum": int(out["game_scores"].sum())}
roots = dev.to_device(empty_states(1))
trees = 148 * per_sm * 2
eng = DeviceUCT(256, n_trees=trees, seed=42, rollout_policy=dev.POLICY_HEURISTIC)
eng.search(roots, 256); torch.cuda.synchronize()
a.record(); r = eng.search(roots, 256); b.record(); torch.cuda.synchronize()
res["trees"] = trees
res["heuristic_rollout_sims_per_sec"] = round(trees * 256 / a.elapsed_time(b) * 1e3)
print(json.dumps(res))

"""ncu driver for the final build: one launch of every kernel family after the 8 corpus-building playout launches.

  ncu --set full --clock-control none --import-source on --launch-skip 8 \
      -k regex:"k_playout|k_movegen|k_compact|k_snapshot|k_replay|k_mcts|k_wide" -o gpurun_out/<name> python tools/profile_all.py
"""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import DeviceUCT
from reinforcementlearning_blokus_b200.state import empty_states

n = 1 << 18
parts = [dev.playout(n // 8, seed=7, game_offset=k * 1_000_000, max_turns=k, want_final=True)["final_states"].clone()
         for k in (6, 12, 18, 24, 30, 36, 42, 48)]                       # 8 skipped launches
corpus = torch.cat(parts).contiguous()
out = {}
dev.playout(65536, seed=42, out=out, want_board_scores=False)             # k_playout<0,false>: the headline launch
tr = dev.playout(65536, seed=42, want_trace=True)["trace"].clone()        # k_playout<0,true>
dev.playout(16384, seed=42, policy=dev.POLICY_HEURISTIC, out={}, want_board_scores=False)   # k_playout<1,...>
dev.movegen_count(corpus)
dev.movegen(corpus)                                                       # count + scan + emit
dev.movegen_bitmap(corpus)
dev.snapshot_features(corpus)
dev.replay_snapshots(tr, (8, 16, 24, 32, 40, 48, 56, 64))
roots = dev.to_device(empty_states(1))
DeviceUCT(256, n_trees=3552, seed=42).search(roots, 256)
# short launches for the two kernels whose counter passes came back as nan in round 1 (185 ms launches under replay)
DeviceUCT(24, n_trees=2960, seed=42, rollout_policy=dev.POLICY_HEURISTIC).search(roots, 24)          # k_mcts<1,false,false>
DeviceUCT(16, n_trees=444, seed=42, leaf_batch=8).search(roots, 16)                                  # k_mcts<0,true,false>
import numpy as np
w = torch.from_numpy(np.random.default_rng(0).normal(0, 0.05, 35))
DeviceUCT(256, n_trees=2368, seed=42, reward_mode=3, leaf_model=w, progressive_bias_weight=0.25).search(roots, 256)   # k_mcts<0,false,true>
# K6: one wide tree (flat bandit and UCT with heuristic rollouts) on a 24-turn position
from reinforcementlearning_blokus_b200.mcts import WideUCT
mid = corpus[3 * (n // 8) + 17: 3 * (n // 8) + 18].contiguous()
WideUCT(1024, width=512, seed=1, max_depth=1, ucb_variant=1, reward_mode=1).search(mid, 1024)        # tables, plan, work<0> x2, finish
WideUCT(512, width=256, seed=1, rollout_policy=dev.POLICY_HEURISTIC).search(mid, 512)                # plan, work<1> x2
torch.cuda.synchronize()
print("ok")

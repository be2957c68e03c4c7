"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck) runs."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.mcts import DeviceUCT, ZobristHash
from reinforcementlearning_blokus_b200.state import empty_states

n = 96
r = dev.playout(n, seed=1, want_trace=True, want_choice=True, want_final=True)
mid = dev.playout(n, seed=2, max_turns=20, want_final=True)["final_states"]
st = torch.cat([r["final_states"], mid, dev.to_device(empty_states(8))]).contiguous()
counts, offsets, ids = dev.movegen(st)
dev.movegen_bitmap(st)
dev.movegen_fixed(st, 64)
dev.has_moves(st)
dev.scores(st)
m = st.shape[0]
first = torch.where(counts > 0, offsets[:-1], torch.zeros_like(offsets[:-1]))
mv = ids.view(torch.int16)[first.long()].view(torch.uint16) if ids.numel() else torch.zeros(m, dtype=torch.uint16, device="cuda")
dev.is_move_legal(st, mv.contiguous())
work = st.clone()
dev.step(work, mv.contiguous(), want_alive=True)
dev.heuristic_scores(st, offsets, ids)
dev.can_place_mask(st, torch.zeros((m, 7), dtype=torch.int64, device="cuda"))
dev.playout(32, seed=3, policy=1, want_trace=True)
dev.playout(32, seed=3, seat_policies=torch.full((32,), 0x00010001, dtype=torch.int32, device="cuda"))
z = ZobristHash(seed=1)
keys = z.hash_states(st)
tt = dev.DeviceTranspositionTable(12)
tt.put(keys, torch.arange(m, dtype=torch.float64, device="cuda"))
tt.get(keys)
for pol in (0, 1):
    eng = DeviceUCT(24, n_trees=4, seed=5, rollout_policy=pol, tt=tt, zobrist=z)
    eng.search(mid[:2].contiguous(), 12)
    eng.search(None, 12, resume=True)
for pol in (0, 1):                                   # leaf-batched rollouts: a CTA of warps per tree
    eng = DeviceUCT(16, n_trees=3, seed=6, rollout_policy=pol, leaf_batch=8 if pol else 16)
    eng.search(mid[:2].contiguous(), 8)
    eng.search(None, 8, resume=True)
eng = DeviceUCT(32, n_trees=3, seed=7, reward_mode=3, leaf_model=torch.linspace(-0.1, 0.1, 35, dtype=torch.float64),
                progressive_bias_weight=0.25, tt=tt, zobrist=z)   # learned leaf evaluation + progressive bias
eng.search(mid[:3].contiguous(), 32)
dev.piece_counts(st)
dev.snapshot_features(st)
dev.replay_snapshots(r["trace"], (0, 8, 40, 64), want_states=True)
torch.cuda.synchronize()
print("sanitize smoke ok", int(counts.sum()))

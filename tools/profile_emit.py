"""ncu driver: the CSR generation kernels on a 2^18-state slice of the SURVEY 8d C1-ii corpus."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.state import empty_states

n = 1 << 18
per = -(-n // 61)
parts = [dev.to_device(empty_states(per))] + [
    dev.playout(per, seed=0xB10C05, game_offset=k * per, max_turns=k, want_final=True)["final_states"].clone() for k in range(1, 61)]
corpus = torch.cat(parts)[:n].contiguous()
for _ in range(2):
    dev.movegen(corpus, one_pass=True)
    dev.movegen_count(corpus)
torch.cuda.synchronize()
print("ok")

"""CSR legal-move lists: one-pass (fixed rows + compaction) against two-pass (count, scan, emit) -- same output, rates."""
import sys
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import device as dev
from reinforcementlearning_blokus_b200.state import empty_states

n = 1 << 20
per = -(-n // 61)
parts = [dev.to_device(empty_states(per))] + [
    dev.playout(per, seed=0xB10C05, game_offset=k * per, max_turns=k, want_final=True)["final_states"].clone() for k in range(1, 61)]
corpus = torch.cat(parts)[:n].contiguous()


def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


one = dev.movegen(corpus, one_pass=True)
two = dev.movegen(corpus, one_pass=False)
same = all(torch.equal(a, b) for a, b in zip(one, two))
dev.ONE_PASS_CAP_SAVED = dev.ONE_PASS_CAP
dev.ONE_PASS_CAP = 256                                   # force the overflow fallback: still the same lists
small = dev.movegen(corpus[:4096].contiguous(), one_pass=True)
ref = dev.movegen(corpus[:4096].contiguous(), one_pass=False)
dev.ONE_PASS_CAP = dev.ONE_PASS_CAP_SAVED
print({"identical": bool(same), "fallback_identical": all(torch.equal(a, b) for a, b in zip(small, ref)),
       "one_pass_G_per_s": round(n / timeit(lambda: dev.movegen(corpus, one_pass=True)) / 1e6, 4),
       "two_pass_G_per_s": round(n / timeit(lambda: dev.movegen(corpus, one_pass=False)) / 1e6, 4)})

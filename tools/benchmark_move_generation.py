#!/usr/bin/env python
"""The reference's own micro-benchmark protocol (benchmarks/benchmark_move_generation.py:136-224) against this
package's drop-in generator: the five seeded benchmark states (generate_random_valid_state(n, seed=n), n = 5, 10, 20,
30, 40 -- stored packed in tests/golden/kat1_states.json), side to move, 50 timed calls each, mean milliseconds.
Also reports the ids-only call (no Python Move objects) and the batched device rate on the same five states.
The reference's numbers for the same states (one Xeon core): 2.36 / 8.01 / 7.22 / 6.31 / 4.94 ms."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reinforcementlearning_blokus_b200 import Board, LegalMoveGenerator  # noqa: E402
from reinforcementlearning_blokus_b200 import device as dev  # noqa: E402
from reinforcementlearning_blokus_b200.state import STATE_DTYPE  # noqa: E402

REFERENCE_MS = {5: 2.36, 10: 8.01, 20: 7.22, 30: 6.31, 40: 4.94}


def main(runs: int = 50):
    with open(os.path.join(ROOT, "tests", "golden", "kat1_states.json")) as f:
        kat = json.load(f)
    gen = LegalMoveGenerator()
    rows = []
    for k in kat:
        st = np.frombuffer(bytes.fromhex(k["packed_hex"]), dtype=STATE_DTYPE)[0]
        board = Board.from_packed(st)
        player = board.current_player
        want = k["legal"][player.value - 1]
        moves = gen.get_legal_moves(board, player)                  # warm-up + correctness
        assert [m.to_id() for m in moves] == want
        t = []
        for _ in range(runs):
            t0 = time.perf_counter(); gen.get_legal_moves(board, player); t.append(time.perf_counter() - t0)
        t_ids = []
        for _ in range(runs):
            t0 = time.perf_counter(); gen.get_legal_move_ids(board, player); t_ids.append(time.perf_counter() - t0)
        rows.append({"plies": k["n"], "legal_moves": len(want), "mean_ms": 1e3 * float(np.mean(t)),
                     "mean_ms_ids_only": 1e3 * float(np.mean(t_ids)), "reference_ms": REFERENCE_MS[k["n"]],
                     "speedup": REFERENCE_MS[k["n"]] / (1e3 * float(np.mean(t)))})
    # the same five states as one device batch, replicated to fill the GPU
    states = np.array([np.frombuffer(bytes.fromhex(k["packed_hex"]), dtype=STATE_DTYPE)[0] for k in kat])
    d = dev.to_device(np.tile(states, 52429)[: 1 << 18])
    dev.movegen(d); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dev.movegen(d); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"protocol": "benchmarks/benchmark_move_generation.py (5 states x %d runs)" % runs, "states": rows,
                      "batched_csr_gens_per_sec": d.shape[0] / e0.elapsed_time(e1) * 1e3}, indent=1))


if __name__ == "__main__":
    main()

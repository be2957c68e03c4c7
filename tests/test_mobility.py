"""Row N4 (mobility block of the snapshot features): host formula vs the reference's own outputs (CPU), device
placement counts vs the reference's (GPU)."""

import os

import numpy as np
import pytest

from reinforcementlearning_blokus_b200 import mobility
from reinforcementlearning_blokus_b200.state import STATE_DTYPE


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    m = np.load(os.path.join(golden_dir, "mobility.npz"))
    return z["states"].view(STATE_DTYPE).reshape(-1), m["piece_counts"], m["metrics"]


def test_formula_matches_reference_bit_for_bit(golden_dir):
    """engine/mobility_metrics.py:38-131 evaluated by the reference (tests/golden/mobility.npz)."""
    states, P, M = _load(golden_dir)
    for i in range(len(states)):
        for p in range(4):
            ref = M[i, p]
            total = int(ref[0])
            amax = int(round(ref[10] * total)) if total else 0
            got = mobility.metrics_from_counts(P[i, p], int(states[i]["used"][p]), amax)
            assert got["totalPlacements"] == total
            assert got["totalOrientationNormalized"] == ref[1] and got["totalCellWeighted"] == ref[2]
            assert [got["buckets"][k] for k in (1, 2, 3, 4, 5)] == ref[3:8].tolist()
            assert got["mobilityEntropy"] == ref[8] and got["pieceTop1Share"] == ref[9]
            assert got["anchorTop1Share"] == ref[10]


@pytest.mark.gpu
def test_device_piece_counts_match_reference(golden_dir):
    torch = pytest.importorskip("torch")
    from reinforcementlearning_blokus_b200 import device as dev

    states, P, M = _load(golden_dir)
    d = dev.to_device(states)
    n = len(states)
    for p in range(4):
        players = torch.full((n,), p, dtype=torch.uint8, device="cuda")
        pc, am = dev.piece_counts(d, players)
        assert np.array_equal(pc.cpu().numpy().astype(np.int32), P[:, p, :])
        total = M[:, p, 0]
        want_am = np.where(total > 0, np.rint(M[:, p, 10] * total), 0).astype(np.int64)
        assert np.array_equal(am.cpu().numpy().astype(np.int64), want_am)
        rows = mobility.batch_mobility(d, players)
        for i in (0, 17, 101, n - 1):
            assert rows[i]["totalCellWeighted"] == M[i, p, 2] and rows[i]["anchorTop1Share"] == M[i, p, 10]

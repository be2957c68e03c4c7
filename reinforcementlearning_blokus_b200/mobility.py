"""Mobility metrics of the reference (engine/mobility_metrics.py:38-131) from device-computed placement counts.

The movegen-bound part -- P_i, the number of legal placements per piece, and the largest number of moves sharing an
anchor cell -- comes from ``bk_movegen_piece_counts`` for a whole batch of (state, player) pairs; the remaining
arithmetic is the reference's formula, evaluated in the same order so the floats are identical:
    PN_i = P_i / O_i,  MW_i = PN_i * S_i,  totals and size buckets summed over unspent pieces in id order,
    entropy = -sum q ln q / ln(#unspent),  pieceTop1Share = max P_i / total,  anchorTop1Share = max anchor / total.
This is the mobility block of the snapshot feature rows (analytics/winprob/features.py:131-193, row N4 of SURVEY 8f);
the geometric features of that file (dead zones, centre proximity, ...) are not movegen-bound and stay outside.
"""

from __future__ import annotations

import math
from typing import Dict, List, Sequence

from . import pieces as P

ORIENTATION_COUNTS = {pid: len(P.PIECE_ORIENTATIONS[pid]) for pid in range(1, 22)}


def metrics_from_counts(piece_counts: Sequence[int], used_mask: int, anchor_max: int) -> Dict[str, object]:
    """One player's PlayerMobilityMetrics (as a dict) from P_i (index pid-1), the spent-piece bit mask and anchor max."""
    total_placements = 0
    total_on = 0.0
    total_cw = 0.0
    buckets = {1: 0.0, 2: 0.0, 3: 0.0, 4: 0.0, 5: 0.0}
    max_piece = 0
    for pid in range(1, 22):
        if (used_mask >> (pid - 1)) & 1:
            continue
        Pi, Oi, Si = int(piece_counts[pid - 1]), ORIENTATION_COUNTS[pid], P.PIECE_SIZES[pid]
        if Pi > max_piece:
            max_piece = Pi
        PNi = Pi / Oi if Oi > 0 else 0.0
        MWi = PNi * Si
        total_placements += Pi
        total_on += PNi
        total_cw += MWi
        buckets[Si] = buckets.get(Si, 0) + MWi
    entropy = 0.0
    if total_placements > 0:
        acc = 0.0
        # the reference sums over its placement dict in first-seen order of its move list, which is ascending piece
        # id (its generator's outer loop runs over pieces): same order, same floats
        for pid in range(1, 22):
            Pi = int(piece_counts[pid - 1])
            if not ((used_mask >> (pid - 1)) & 1) and Pi > 0:
                q = Pi / total_placements
                acc -= q * math.log(q)
        avail = 21 - bin(used_mask & 0x1FFFFF).count("1")
        entropy = acc / math.log(avail) if avail > 1 else 0.0
    return {
        "totalPlacements": total_placements, "totalOrientationNormalized": total_on, "totalCellWeighted": total_cw,
        "buckets": buckets, "mobilityEntropy": entropy,
        "pieceTop1Share": max_piece / total_placements if total_placements > 0 else 0.0,
        "anchorTop1Share": int(anchor_max) / total_placements if total_placements > 0 else 0.0,
    }


def batch_mobility(states, players=None) -> List[Dict[str, object]]:
    """Mobility metrics of every (state, player) pair of a device batch: one kernel launch plus host arithmetic."""
    from . import device as dev

    pc, am = dev.piece_counts(states, players)
    pc_h, am_h = pc.cpu().numpy(), am.cpu().numpy()
    raw = states.cpu().numpy()
    used = raw[:, 224:240].copy().view("<u4").reshape(-1, 4)
    who = players.cpu().numpy() if players is not None else raw[:, 240]
    return [metrics_from_counts(pc_h[i], int(used[i, int(who[i]) & 3]), int(am_h[i])) for i in range(len(raw))]

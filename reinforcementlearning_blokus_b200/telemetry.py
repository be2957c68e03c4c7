"""Per-player telemetry of the reference (engine/telemetry.py:17-198) with the movegen-bound part as device batches.

``collect_all_player_metrics`` is what makes the reference's ``make_move`` 170x slower when telemetry is on (SURVEY.md
Q11): per player one legal-move list, the mobility block, and ``simulate_mobility_stability`` -- for each of the three
opponents up to k sampled moves, each played on a copy of the board and followed by a full legal-move generation for
the player.  Here the whole board costs FOUR launches regardless of k:

    1. ``bk_movegen_*``               legal lists of all four players (ascending ids)
    2. ``bk_movegen_piece_counts``    P_i per piece + anchor maximum of all four players (mobility block)
    3. ``bk_step``                    every sampled (opponent, move) of every player on its own copy (<= 4 x 3 x 8 states)
    4. ``bk_movegen_count``           the player's mobility on every copy

The sampling itself is the reference's: one ``random.Random(TelemetryConfig.DETERMINISM_SEED)`` per player, one
``shuffle`` of each opponent's list, the first k entries (engine/telemetry.py:23-46).  ``shuffle`` depends on the
ORDER of the list, and the reference's lists come in CPython-set order, so sampled moves are comparable
only when both sides list moves in the same order; the golden fixture (tests/golden/generate_golden.py gen_telemetry)
runs the reference with its generator wrapped to ascending ids -- the order of this package.

The geometric metrics (engine/advanced_metrics.py: frontier spread, effective frontier, centre proximity, dead-space
split) are not movegen-bound; they are evaluated on the host from the board's bitboards.
"""

from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import pieces as P
from .board import Board, Player

_PLAYERS = list(Player)


class TelemetryConfig:
    """engine/metrics_config.py:8-73: sampling knobs, polarity of every metric (+1 = higher is better) and the weights
    of the win-proxy score."""

    MOBILITY_SAMPLES_K = 3
    ANCHOR_RADIUS = 3
    DETERMINISM_SEED = 42
    METRIC_POLARITY: Dict[str, int] = {
        "frontierSize": 1, "effectiveFrontier": 1, "frontierComponentCount": 1, "frontierQuadrantCoverage": 1,
        "mobility": 1, "mobilityWeighted": 1, "mobilityEntropy": 1, "pieceTop1Share": -1, "anchorTop1Share": -1,
        "mobilityNextMean": 1, "mobilityNextP10": 1, "mobilityNextMin": 1, "mobilityDropRisk": -1,
        "centerControl": 1, "centerControlWeighted": 1,
        "deadSpaceNearOpponents": 1, "deadSpaceNearSelf": -1,
        "lockedArea": -1, "criticalPiecesCount": -1, "bottleneckScore": 1, "remainingArea": -1,
        "largestRemainingPiece": -1, "unloadPotential": 1,
        "deadSpace": 1, "pieceLockRisk": -1,
    }
    WIN_PROXY_WEIGHTS: Dict[str, float] = {
        "remainingAreaAdv": 2.0, "effectiveFrontierAdv": 1.5, "mobilityNextP10Adv": 1.5, "lockedAreaAdv": 1.0,
        "centerControlWeightedAdv": 0.5,
    }


# ---------------------------------------------------------------------------------------------------------------------
# device batches
# ---------------------------------------------------------------------------------------------------------------------
def _legal_lists(board: Board) -> List[np.ndarray]:
    """Ascending legal-move ids of all four players: one generation launch over four (state, player) pairs."""
    import torch

    from . import device as dev

    st = dev.to_device(np.atleast_1d(board.to_packed())).repeat(4, 1).contiguous()
    players = torch.arange(4, dtype=torch.uint8, device=st.device)
    counts, offsets, ids = dev.movegen(st, players)
    off = offsets.cpu().numpy().astype(np.int64)
    flat = ids.cpu().numpy()
    return [flat[off[k]:off[k + 1]].copy() for k in range(4)]


def _sample_plan(lists: Sequence[np.ndarray], player_idx: int, k_samples: int) -> List[Tuple[int, int]]:
    """(opponent index, move id) pairs simulate_mobility_stability evaluates for one player, in its order."""
    rng = random.Random(TelemetryConfig.DETERMINISM_SEED)
    plan = []
    for opp in range(4):
        if opp == player_idx:
            continue
        moves = [int(m) for m in lists[opp]]
        if not moves:
            continue
        rng.shuffle(moves)                      # engine/telemetry.py:45 -- same list length and order => same samples
        plan.extend((opp, m) for m in moves[:k_samples])
    return plan


def _mobility_after(board: Board, plans: Sequence[Sequence[Tuple[int, int]]]) -> List[List[int]]:
    """plans[p] = (mover, move id) samples of player p; returns the player's legal-move count after each sample.
    One bk_step + one bk_movegen_count over all samples of all players."""
    import torch

    from . import device as dev

    flat = [(p, mover, mid) for p, plan in enumerate(plans) for mover, mid in plan]
    if not flat:
        return [[] for _ in plans]
    n = len(flat)
    st = dev.to_device(np.atleast_1d(board.to_packed())).repeat(n, 1).contiguous()
    movers = torch.tensor([f[1] for f in flat], dtype=torch.uint8, device=st.device)
    who = torch.tensor([f[0] for f in flat], dtype=torch.uint8, device=st.device)
    ids = torch.from_numpy(np.array([f[2] for f in flat], dtype=np.uint16).view(np.int16)).view(torch.uint16).to(st.device)
    ok, _ = dev.step(st, ids, players=movers)
    counts = dev.movegen_count(st, who)
    got = torch.stack([ok.to(torch.int32), counts]).cpu().numpy()
    if not got[0].all():
        raise RuntimeError("a sampled legal move was rejected by bk_step")     # cannot happen: the ids come from movegen
    out: List[List[int]] = [[] for _ in plans]
    for (p, _, _), c in zip(flat, got[1]):
        out[p].append(int(c))
    return out


def _stability_from_samples(current_mobility: int, sampled: Sequence[int]) -> Dict[str, float]:
    """engine/telemetry.py:28-77 after the sampling: mean / P10 / min / drop risk."""
    if current_mobility == 0:
        return {"mobilityNextMean": 0.0, "mobilityNextP10": 0.0, "mobilityNextMin": 0.0, "mobilityDropRisk": 0.0}
    if not sampled:
        c = float(current_mobility)
        return {"mobilityNextMean": c, "mobilityNextP10": c, "mobilityNextMin": c, "mobilityDropRisk": 0.0}
    vals = sorted(float(v) for v in sampled)
    n = len(vals)
    p10 = vals[max(0, int(0.10 * n) - 1)]
    return {"mobilityNextMean": sum(vals) / n, "mobilityNextP10": p10, "mobilityNextMin": vals[0],
            "mobilityDropRisk": max(0.0, current_mobility - p10)}


def simulate_mobility_stability(board: Board, player: Player, move_generator=None, fast_mode: bool = True) -> Dict[str, float]:
    """engine/telemetry.py:17-77.  ``move_generator`` is accepted for signature compatibility; the lists come from
    the device generator (ascending ids)."""
    k = TelemetryConfig.MOBILITY_SAMPLES_K if fast_mode else 8
    lists = _legal_lists(board)
    pi = player.value - 1
    if len(lists[pi]) == 0:
        return _stability_from_samples(0, [])
    plan = _sample_plan(lists, pi, k)
    plans = [plan if q == pi else [] for q in range(4)]
    return _stability_from_samples(len(lists[pi]), _mobility_after(board, plans)[pi])


# ---------------------------------------------------------------------------------------------------------------------
# geometric metrics (engine/advanced_metrics.py), host side
# ---------------------------------------------------------------------------------------------------------------------
def _frontier_cells(board: Board, player: Player) -> List[Tuple[int, int]]:
    m = board.player_frontier_bits[player]
    return [(i // 20, i % 20) for i in range(400) if (m >> i) & 1]


def compute_center_proximity(board: Board, player: Player) -> int:
    """engine/advanced_metrics.py:71-86: smallest int(|r-9.5|-0.5 clipped + |c-9.5|-0.5 clipped) over own cells, 9 if none."""
    rows, cols = np.nonzero(board.grid == player.value)
    if rows.size == 0:
        return 9
    d = np.maximum(0.0, np.abs(rows - 9.5) - 0.5) + np.maximum(0.0, np.abs(cols - 9.5) - 0.5)
    return min(9, int(d.min()))


def compute_frontier_spread(board: Board, player: Player) -> Tuple[int, int]:
    """engine/advanced_metrics.py:286-330: (clusters of frontier cells linked at Chebyshev distance <= 2, quadrants hit)."""
    cells = _frontier_cells(board, player)
    if not cells:
        return 0, 0
    pts = np.array(cells)
    n = len(cells)
    link = (np.abs(pts[:, None, 0] - pts[None, :, 0]) <= 2) & (np.abs(pts[:, None, 1] - pts[None, :, 1]) <= 2)
    label = -np.ones(n, dtype=int)
    comps = 0
    for s in range(n):
        if label[s] >= 0:
            continue
        stack = [s]
        label[s] = comps
        while stack:
            i = stack.pop()
            for j in np.nonzero(link[i] & (label < 0))[0]:
                label[j] = comps
                stack.append(int(j))
        comps += 1
    quads = {(r >= 10, c >= 10) for r, c in cells}
    return comps, len(quads)


def compute_effective_frontier(board: Board, player: Player) -> float:
    """engine/advanced_metrics.py:218-283: sum over frontier cells of log1p(space) * (1 - min(0.8, vulnerability)), space =
    empty cells of the radius-3 box without an own edge neighbour, vulnerability = 0.2 per opponent frontier cell in the
    box (added one by one, as the reference does).  The reference accumulates over a Python set; the sum here runs in
    ascending cell order, so the two can differ in the last bits (the per-cell weights are identical)."""
    cells = _frontier_cells(board, player)
    if not cells:
        return 0.0
    R = TelemetryConfig.ANCHOR_RADIUS
    grid = board.grid
    own = grid == player.value
    orth = np.zeros_like(own)
    orth[1:, :] |= own[:-1, :]
    orth[:-1, :] |= own[1:, :]
    orth[:, 1:] |= own[:, :-1]
    orth[:, :-1] |= own[:, 1:]
    space = (grid == 0) & ~orth
    opp = np.zeros((20, 20), dtype=bool)
    for q in _PLAYERS:
        if q != player:
            for r, c in _frontier_cells(board, q):
                opp[r, c] = True
    total = 0.0
    for r, c in cells:
        r0, r1, c0, c1 = max(0, r - R), min(19, r + R), max(0, c - R), min(19, c + R)
        space_score = int(space[r0:r1 + 1, c0:c1 + 1].sum())
        vuln = 0.0
        for _ in range(int(opp[r0:r1 + 1, c0:c1 + 1].sum())):
            vuln += 0.2
        total += math.log1p(space_score) * (1.0 - min(0.8, vuln))
    return total


def compute_dead_space_split(board: Board, player: Player) -> Tuple[float, float]:
    """engine/advanced_metrics.py:108-175: empty regions (4-connected) holding no frontier cell of anybody are dead; a
    dead region counts as near-self if it borders a cell of `player`, else near-opponents if it borders anyone."""
    from scipy import ndimage

    grid = board.grid
    labels, n = ndimage.label(grid == 0)
    if n == 0:
        return 0.0, 0.0
    alive = np.zeros(n + 1, dtype=bool)
    for q in _PLAYERS:
        for r, c in _frontier_cells(board, q):
            alive[labels[r, c]] = True
    alive[0] = True
    own = grid == player.value
    others = (grid != 0) & ~own

    def touched(mask):
        t = np.zeros_like(mask)
        t[1:, :] |= mask[:-1, :]
        t[:-1, :] |= mask[1:, :]
        t[:, 1:] |= mask[:, :-1]
        t[:, :-1] |= mask[:, 1:]
        return t

    near_self = np.zeros(n + 1, dtype=bool)
    near_opp = np.zeros(n + 1, dtype=bool)
    near_self[np.unique(labels[touched(own) & (labels > 0)])] = True
    near_opp[np.unique(labels[touched(others) & (labels > 0)])] = True
    sizes = np.bincount(labels.ravel(), minlength=n + 1)
    dead = ~alive
    ds_self = int(sizes[dead & near_self].sum())
    ds_opp = int(sizes[dead & ~near_self & near_opp].sum())
    return float(ds_self), float(ds_opp)


# ---------------------------------------------------------------------------------------------------------------------
# engine/telemetry.py:79-198
# ---------------------------------------------------------------------------------------------------------------------
def _static_block(board: Board, player: Player) -> Dict[str, float]:
    comp_count, quad_cov = compute_frontier_spread(board, player)
    used = board.player_pieces_used[player]
    remaining_area = sum(P.PIECE_SIZES[pid] for pid in range(1, 22) if pid not in used)
    phase_ratio = max(0.0, min(1.0, 1.0 - (remaining_area / 89.0)))
    center_raw = 9.0 - compute_center_proximity(board, player)
    return {
        "frontierSize": float(len(_frontier_cells(board, player))),
        "effectiveFrontier": float(compute_effective_frontier(board, player)),
        "frontierComponentCount": float(comp_count),
        "frontierQuadrantCoverage": float(quad_cov),
        "centerControl": float(center_raw),
        "centerControlWeighted": float(center_raw * phase_ratio),
        "pieceLockRisk": 0.0,
        "remainingArea": float(remaining_area),
    }


def _movegen_block(board: Board, player: Player, piece_counts: Sequence[int], anchor_max: int,
                   stability: Dict[str, float]) -> Dict[str, float]:
    """The generator-dependent entries of compute_player_metrics (engine/telemetry.py:112-163) from P_i per piece."""
    from .mobility import metrics_from_counts

    used = board.player_pieces_used[player]
    used_mask = sum(1 << (pid - 1) for pid in used if 1 <= pid <= 21)
    unspent = [pid for pid in range(1, 22) if pid not in used]
    counts = {pid: int(piece_counts[pid - 1]) for pid in unspent}
    mob = metrics_from_counts(piece_counts, used_mask, anchor_max)
    out = {"mobility": float(sum(counts.values())), "mobilityWeighted": mob["totalCellWeighted"],
           "mobilityEntropy": mob["mobilityEntropy"], "pieceTop1Share": mob["pieceTop1Share"],
           "anchorTop1Share": mob["anchorTop1Share"]}
    out.update(stability)
    locked = [pid for pid in unspent if counts[pid] == 0]
    out["pieceLockRisk"] = float(len(locked))
    out["lockedArea"] = float(sum(P.PIECE_SIZES[pid] for pid in locked))
    out["criticalPiecesCount"] = float(sum(1 for c in counts.values() if 0 < c <= 3))
    placed = [c for c in counts.values() if c > 0]
    out["bottleneckScore"] = float(min(placed)) if placed else 0.0
    out["largestRemainingPiece"] = float(max((P.PIECE_SIZES[pid] for pid in unspent), default=0))
    out["unloadPotential"] = float(sum(P.PIECE_SIZES[pid] for pid in unspent if counts[pid] > 0))
    return out


def _fallback_block(frontier_size: float) -> Dict[str, float]:
    mob = float(frontier_size * 2)                     # engine/telemetry.py:165-180: no generator available
    return {"mobility": mob, "mobilityWeighted": mob * 3, "mobilityEntropy": 0.0, "pieceTop1Share": 0.0,
            "anchorTop1Share": 0.0, "mobilityNextMean": mob, "mobilityNextP10": mob, "mobilityNextMin": mob,
            "mobilityDropRisk": 0.0, "lockedArea": 0.0, "criticalPiecesCount": 0.0, "bottleneckScore": 0.0,
            "largestRemainingPiece": 0.0, "unloadPotential": 0.0}


def _device_blocks(board: Board, players: Sequence[Player], fast_mode: bool) -> Dict[Player, Dict[str, float]]:
    """The generator-dependent metrics of the given players: four launches for the whole board."""
    import torch

    from . import device as dev

    k = TelemetryConfig.MOBILITY_SAMPLES_K if fast_mode else 8
    lists = _legal_lists(board)
    st = dev.to_device(np.atleast_1d(board.to_packed())).repeat(4, 1).contiguous()
    pc, am = dev.piece_counts(st, torch.arange(4, dtype=torch.uint8, device=st.device))
    pc_h, am_h = pc.cpu().numpy(), am.cpu().numpy()
    want = {p.value - 1 for p in players}
    plans = [(_sample_plan(lists, q, k) if (q in want and len(lists[q])) else []) for q in range(4)]
    after = _mobility_after(board, plans)
    out = {}
    for p in players:
        q = p.value - 1
        stab = _stability_from_samples(len(lists[q]), after[q])
        out[p] = _movegen_block(board, p, pc_h[q], int(am_h[q]), stab)
    return out


def compute_player_metrics(board: Board, player: Player, move_generator=None, fast_mode: bool = True) -> Dict[str, float]:
    """engine/telemetry.py:79-182."""
    metrics = _static_block(board, player)
    if move_generator is not None:
        metrics.update(_device_blocks(board, [player], fast_mode)[player])
    else:
        metrics.update(_fallback_block(metrics["frontierSize"]))
    return metrics


def collect_all_player_metrics(board: Board, move_generator=None, fast_mode: bool = True) -> Dict[str, Dict[str, float]]:
    """engine/telemetry.py:184-198: all four players, the device work shared (four launches per board)."""
    blocks = _device_blocks(board, _PLAYERS, fast_mode) if move_generator is not None else None
    out: Dict[str, Dict[str, float]] = {}
    for p in _PLAYERS:
        m = _static_block(board, p)
        m.update(blocks[p] if blocks is not None else _fallback_block(m["frontierSize"]))
        ds_self, ds_opp = compute_dead_space_split(board, p)
        m["deadSpaceNearSelf"] = ds_self
        m["deadSpaceNearOpponents"] = ds_opp
        m["deadSpace"] = ds_self + ds_opp
        out[p.name] = m
    return out


def _advantage(state: Dict[str, Dict[str, float]], who: str, key: str) -> float:
    """polarity * (own value - mean of the opponents' values), engine/telemetry.py:234-242."""
    if who not in state or key not in state[who]:
        return 0.0
    others = [state[q][key] for q in state if q != who and key in state[q]]
    mean = sum(others) / len(others) if others else 0.0
    return float(TelemetryConfig.METRIC_POLARITY.get(key, 1) * (state[who][key] - mean))


def _win_proxy(state: Dict[str, Dict[str, float]], who: str) -> float:
    score = 0.0
    for name, weight in TelemetryConfig.WIN_PROXY_WEIGHTS.items():
        key = name.replace("Adv", "")
        if who in state and key in state[who]:
            score += _advantage(state, who, key) * weight
    return score


def compute_move_telemetry_delta(ply: int, mover_id: str, move_id: str, before: Dict[str, Dict[str, float]],
                                 after: Dict[str, Dict[str, float]]) -> Dict[str, object]:
    """engine/telemetry.py:200-276: what one move changed -- for the mover, for each opponent and in total, as
    advantage over the opponents' mean, and in the weighted win-proxy score."""
    if mover_id not in before or mover_id not in after:
        return {}
    delta_self = {k: v - before[mover_id].get(k, 0.0) for k, v in after[mover_id].items()}
    by_player: Dict[str, Dict[str, float]] = {}
    total: Dict[str, float] = {}
    for name, metrics in after.items():
        if name == mover_id:
            continue
        was = before.get(name, {})
        by_player[name] = {}
        for k, v in metrics.items():
            d = v - was.get(k, 0.0)
            by_player[name][k] = d
            total[k] = total.get(k, 0.0) + d
    adv = {k + "Adv": _advantage(after, mover_id, k) - _advantage(before, mover_id, k)
           for k in TelemetryConfig.METRIC_POLARITY if k in after[mover_id]}
    adv["winProxy"] = _win_proxy(after, mover_id) - _win_proxy(before, mover_id)
    return {"ply": ply, "moverId": mover_id, "moveId": move_id, "deltaSelf": delta_self, "deltaOppTotal": total,
            "deltaOppByPlayer": by_player, "deltaAdvantage": adv, "winProxyDelta": adv["winProxy"]}

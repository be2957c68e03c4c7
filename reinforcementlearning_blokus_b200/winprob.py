"""Win-probability snapshot features and the learned leaf evaluator.

Mirrors ``analytics/winprob/features.py`` (``SNAPSHOT_FEATURE_COLUMNS``, ``build_snapshot_runtime_context``,
``extract_player_snapshot_features``, ``coerce_feature_dict``), the snapshot rows of
``analytics/tournament/arena_runner.py:522-575,722-731,815-858`` and ``mcts/learned_evaluator.py``.  All 34 columns of all
four players of a board come from ONE device call (``bk_snapshot_features``: four piece-count move generations, a flood
fill for the dead zones, popcounts for the rest); whole datasets come from ``bk_replay_snapshots``, which replays
recorded games on the device and captures the rows at the checkpoint plies.  Values are bit-identical to the
reference's (``tests/golden/snapshots.npz``).
"""

from __future__ import annotations

import json
import math
from collections.abc import Mapping
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .board import Board, Player

SNAPSHOT_FEATURE_COLUMNS: List[str] = [          # analytics/winprob/features.py:23-58, in the order of the device row
    "frontier_size", "corner_differential", "center_proximity", "opponent_adjacency", "deadzone_count",
    "deadzone_fraction", "mobility_total_placements", "mobility_total_orientation_normalized",
    "mobility_total_cell_weighted", "mobility_bucket_1", "mobility_bucket_2", "mobility_bucket_3", "mobility_bucket_4",
    "mobility_bucket_5", "blocking_exposure_legal_moves_opp_sum", "utility_frontier_plus_mobility", "pieces_used_count",
    "pieces_remaining_count", "remaining_squares", "remaining_size_1_count", "remaining_size_2_count",
    "remaining_size_3_count", "remaining_size_4_count", "remaining_size_5_count", "remaining_key_piece_17",
    "remaining_key_piece_19", "remaining_key_piece_20", "phase_ply", "phase_turn_index", "phase_board_occupancy",
    "phase_piece_usage_ratio", "phase_progress_turn_ratio", "phase_progress_placement_ratio", "player_board_occupancy",
]
DEFAULT_SNAPSHOT_PLYS = [8, 16, 24, 32, 40, 48, 56, 64]          # arena_runner.py:44


def board_feature_block(board: Board, turn_index: Optional[int] = None, max_turns: int = 2500) -> np.ndarray:
    """float64[4, 34]: the feature rows of the four players of one board (one upload, one kernel, one download)."""
    import torch

    from . import device as dev

    st = dev.to_device(np.atleast_1d(board.to_packed()))
    fr = torch.from_numpy(board.frontier_words()[None].copy()).to(st.device)
    ti = torch.tensor([board.move_count if turn_index is None else int(turn_index)], dtype=torch.int32, device=st.device)
    return dev.snapshot_features(st, fr, ti, max_turns)[0].cpu().numpy()


@dataclass(frozen=True)
class SnapshotRuntimeContext:
    """features.py:61-71, plus the device block shared by the four per-player extractions of one snapshot."""

    ply: int
    turn_index: int
    board_occupancy: float
    deadzone_count: int
    deadzone_fraction: float
    progress_turn_ratio: float
    progress_placement_ratio: float
    block: Any = None


def build_snapshot_runtime_context(board: Board, *, turn_index: int, max_turns: int) -> SnapshotRuntimeContext:
    blk = board_feature_block(board, turn_index, max_turns)
    col = SNAPSHOT_FEATURE_COLUMNS.index
    return SnapshotRuntimeContext(
        ply=int(board.move_count), turn_index=int(turn_index), board_occupancy=float(blk[0, col("phase_board_occupancy")]),
        deadzone_count=int(blk[0, col("deadzone_count")]), deadzone_fraction=float(blk[0, col("deadzone_fraction")]),
        progress_turn_ratio=float(blk[0, col("phase_progress_turn_ratio")]),
        progress_placement_ratio=float(blk[0, col("phase_progress_placement_ratio")]), block=blk)


def extract_player_snapshot_features(board: Board, *, player: Player, context: SnapshotRuntimeContext,
                                     move_generator=None) -> Dict[str, float]:
    """features.py:131-193; `move_generator` is accepted for signature compatibility (the device generates)."""
    blk = context.block if context.block is not None else board_feature_block(board, context.turn_index, 2500)
    row = blk[player.value - 1]
    return {name: float(row[i]) for i, name in enumerate(SNAPSHOT_FEATURE_COLUMNS)}


def coerce_feature_dict(features: Mapping[str, Any]) -> Dict[str, float]:
    return {c: float(features.get(c, 0.0)) for c in SNAPSHOT_FEATURE_COLUMNS}


# ---- dataset rows (arena_runner.py:522-575, 722-731) ---------------------------------------------------------
def snapshot_rows_from_replay(replay: Mapping[str, Any], records: Sequence[Mapping[str, Any]]) -> List[Dict[str, Any]]:
    """Rows of snapshots.csv / snapshots.parquet for a batch of replayed games.

    replay: the dict of ``device.replay_snapshots`` (features [N,K,4,34], turn_index [N,K], checkpoints [K]);
    records: the N game records of ``arena.run_device_arena`` (run_id, game_id, seeds, seats, final scores / ranks).
    Row keys and order follow ``_build_snapshot_rows_for_checkpoint`` plus the labels filled in at the end of the game."""
    feats = replay["features"].cpu().numpy()
    turns = replay["turn_index"].cpu().numpy()
    cks = [int(c) for c in replay["checkpoints"]]
    rows: List[Dict[str, Any]] = []
    for g, rec in enumerate(records):
        scores = rec["final_scores"]
        winner_ids = list(rec["winner_ids"])
        for k, ply in enumerate(cks):
            if turns[g, k] < 0:
                continue
            # the side to move after `ply` placements and the passes in between: turn_index turns from RED
            current = int(turns[g, k]) % 4 + 1
            for p in Player:
                pid = int(p.value)
                row: Dict[str, Any] = {
                    "run_id": rec["run_id"], "game_id": rec["game_id"], "game_index": int(rec["game_index"]),
                    "game_seed": int(rec["game_seed"]), "checkpoint_index": k, "checkpoint_ply": ply, "ply": ply,
                    "turn_index": int(turns[g, k]), "current_player_id": current, "player_id": pid, "player_name": p.name,
                    "agent_name": rec["seat_assignment"][str(pid)], "seat_index": pid - 1,
                    "winner_id": rec["winner_id"], "winner_ids_json": json.dumps(winner_ids, sort_keys=True),
                    "label_is_winner": int(pid in winner_ids), "final_score": int(scores[str(pid)]),
                    "final_rank": int(rec["final_ranks"][str(pid)]),
                    "final_scores_json": json.dumps(scores, sort_keys=True), "is_tie": bool(rec["is_tie"]),
                }
                row.update({c: float(feats[g, k, pid - 1, i]) for i, c in enumerate(SNAPSHOT_FEATURE_COLUMNS)})
                rows.append(row)
    return rows


def write_snapshots_dataset(run_dir: str, snapshot_rows: Sequence[Mapping[str, Any]]) -> Dict[str, Any]:
    """arena_runner.py:815-858: snapshots.csv always, snapshots.parquet when a parquet engine is present."""
    import os

    if not snapshot_rows:
        return {"enabled": False, "rows": 0, "path_parquet": None, "path_csv": None, "parquet_written": False,
                "parquet_error": None}
    import pandas as pd

    frame = pd.DataFrame(list(snapshot_rows))
    csv_path = os.path.join(run_dir, "snapshots.csv")
    frame.to_csv(csv_path, index=False)
    parquet_path, written, err = os.path.join(run_dir, "snapshots.parquet"), False, None
    try:
        frame.to_parquet(parquet_path, index=False)
        written = True
    except Exception as exc:                                  # depends on the parquet engine, as in the reference
        err = str(exc)
    return {"enabled": True, "rows": int(len(frame)), "path_parquet": parquet_path if written else None,
            "path_csv": csv_path, "parquet_written": written, "parquet_error": err}


def fold_logistic_pipeline(artifact: Mapping[str, Any]) -> Optional[np.ndarray]:
    """float64[35] (34 weights in SNAPSHOT_FEATURE_COLUMNS order + bias) of a "pairwise_logreg" artifact whose pipeline is
    [StandardScaler,] LogisticRegression over (a subset of) the snapshot columns -- the form the device kernel evaluates:
    logit = intercept + sum_c coef_c * (d_c - mean_c) / scale_c.  None when the pipeline has any other shape."""
    if artifact.get("model_type") != "pairwise_logreg":
        return None
    pipe = artifact.get("pipeline")
    steps = [st for _, st in getattr(pipe, "steps", [])] if hasattr(pipe, "steps") else [pipe]
    if not steps or type(steps[-1]).__name__ != "LogisticRegression" or len(steps) > 2:
        return None
    clf = steps[-1]
    coef = np.asarray(clf.coef_, dtype=np.float64).reshape(-1)
    cols = list(artifact.get("feature_columns", SNAPSHOT_FEATURE_COLUMNS))
    if len(coef) != len(cols) or any(c not in SNAPSHOT_FEATURE_COLUMNS for c in cols) or len(np.atleast_1d(clf.intercept_)) != 1:
        return None
    mean, scale = np.zeros(len(cols)), np.ones(len(cols))
    if len(steps) == 2:
        sc = steps[0]
        if type(sc).__name__ != "StandardScaler":
            return None
        if getattr(sc, "with_mean", True) and getattr(sc, "mean_", None) is not None:
            mean = np.asarray(sc.mean_, dtype=np.float64)
        if getattr(sc, "with_std", True) and getattr(sc, "scale_", None) is not None:
            scale = np.asarray(sc.scale_, dtype=np.float64)
    w = coef / scale
    out = np.zeros(35, dtype=np.float64)
    for c, v in zip(cols, w):
        out[SNAPSHOT_FEATURE_COLUMNS.index(c)] = v
    out[34] = float(np.atleast_1d(clf.intercept_)[0]) - float(np.sum(w * mean))
    return out


def flatten_gbt_phase_models(artifact: Mapping[str, Any]) -> Optional[np.ndarray]:
    """One float64 blob for the UCT kernel from a "pairwise_gbt_phase" artifact (mcts/learned_evaluator.py:112-127) whose
    phase / fallback models are scikit-learn GradientBoostingClassifier (binary) over all 34 snapshot columns:

        blob[0:3]            offsets of the early / mid / late model (a missing phase points at the fallback model)
        model: n_trees, init_raw, learning_rate, 0, roots[n_trees], nodes[*][5] = (feature or -1, threshold, left, right, value)

    raw = init_raw + learning_rate * sum_t value(leaf_t(x)), p = sigmoid(raw); a tree goes left when float32(x[feature])
    <= threshold, as scikit-learn's trees do.  None when the artifact has any other shape."""
    if artifact.get("model_type") != "pairwise_gbt_phase":
        return None
    cols = list(artifact.get("feature_columns", SNAPSHOT_FEATURE_COLUMNS))
    if cols != SNAPSHOT_FEATURE_COLUMNS:
        return None
    phases = artifact.get("phase_models", {}) or {}
    fallback = artifact.get("fallback_model")
    chosen = [phases.get(ph) or fallback for ph in ("early", "mid", "late")]
    if any(m is None for m in chosen):
        return None
    parts, offsets, seen = [], [], {}
    pos = 3
    for m in chosen:
        if id(m) in seen:
            offsets.append(seen[id(m)])
            continue
        if type(m).__name__ != "GradientBoostingClassifier" or getattr(m, "n_classes_", 2) != 2 or m.estimators_.shape[1] != 1:
            return None
        init_raw = float(np.asarray(m._raw_predict_init(np.zeros((1, len(cols))))).ravel()[0])
        roots, nodes, base = [], [], 0
        for est in m.estimators_[:, 0]:
            t = est.tree_
            roots.append(float(base))
            for k in range(t.node_count):
                leaf = t.children_left[k] < 0
                nodes.append([-1.0 if leaf else float(t.feature[k]), float(t.threshold[k]),
                              -1.0 if leaf else float(base + t.children_left[k]), -1.0 if leaf else float(base + t.children_right[k]),
                              float(t.value[k, 0, 0])])
            base += t.node_count
        block = np.concatenate([[float(len(roots)), init_raw, float(m.learning_rate), 0.0], roots, np.asarray(nodes).ravel()])
        seen[id(m)] = pos
        offsets.append(pos)
        parts.append(block)
        pos += len(block)
    return np.concatenate([np.asarray(offsets, dtype=np.float64)] + parts).astype(np.float64)


# ---- learned leaf evaluator (mcts/learned_evaluator.py) ----------------------------------------------------------
def _phase_bucket_from_occupancy(board_occupancy: float, *, early_max: float = 0.33, mid_max: float = 0.66) -> str:
    return "early" if board_occupancy < early_max else ("mid" if board_occupancy < mid_max else "late")


class LearnedWinProbabilityEvaluator:
    """Same constructor, artifact format and methods as mcts/learned_evaluator.py:36-152.  The feature rows come from
    the device; the pairwise model (a scikit-learn pipeline or per-phase gradient-boosted models stored with joblib) is
    evaluated by scikit-learn on the host, for all three opponents in one ``predict_proba`` call."""

    def __init__(self, artifact_path: str, *, max_turns: int = 2500, potential_mode: str = "prob") -> None:
        self.artifact_path = str(artifact_path)
        self.max_turns = int(max_turns)
        self.potential_mode = potential_mode
        self._is_dummy = self.artifact_path.endswith("dummy_model.json")
        if self._is_dummy:
            self.model_type, self.feature_columns = "dummy", []
            return
        import joblib

        self.artifact: Dict[str, Any] = joblib.load(self.artifact_path)
        self.model_type = str(self.artifact.get("model_type", ""))
        self.feature_columns = list(self.artifact.get("feature_columns", SNAPSHOT_FEATURE_COLUMNS))
        if self.model_type not in {"pairwise_logreg", "pairwise_gbt_phase"}:
            raise ValueError(f"Unsupported learned evaluator model_type '{self.model_type}' in {self.artifact_path}.")
        self._cols = [SNAPSHOT_FEATURE_COLUMNS.index(c) for c in self.feature_columns]
        # the model in the form the UCT kernel evaluates: float64[35] (logistic) or a flattened tree blob (GBT), or None
        self.device_model = fold_logistic_pipeline(self.artifact)
        self.device_model_kind = 0
        if self.device_model is None:
            self.device_model = flatten_gbt_phase_models(self.artifact)
            self.device_model_kind = 1

    def _extract_features_for_all_players(self, board: Board) -> Dict[int, Dict[str, float]]:
        blk = board_feature_block(board, int(board.move_count), self.max_turns)
        return {p + 1: {c: float(blk[p, i]) for i, c in enumerate(SNAPSHOT_FEATURE_COLUMNS)} for p in range(4)}

    def _pairwise_matrix(self, blk: np.ndarray, x: np.ndarray, i: int, others: Sequence[int]) -> np.ndarray:
        if self.model_type == "pairwise_logreg":
            return self.artifact["pipeline"].predict_proba(x)[:, 1]
        occ_col = SNAPSHOT_FEATURE_COLUMNS.index("phase_board_occupancy")
        out = np.zeros(len(others))
        for n, j in enumerate(others):
            phase = _phase_bucket_from_occupancy(float((blk[i, occ_col] + blk[j, occ_col]) / 2.0))
            model = self.artifact.get("phase_models", {}).get(phase) or self.artifact.get("fallback_model")
            if model is None:
                raise RuntimeError("GBT evaluator artifact missing both phase model and fallback model.")
            out[n] = model.predict_proba(x[n:n + 1])[:, 1][0]
        return out

    def predict_player_win_probability(self, board: Board, player: Player) -> float:
        if self._is_dummy:
            return 0.5
        blk = board_feature_block(board, int(board.move_count), self.max_turns)
        i = player.value - 1
        others = [j for j in range(4) if j != i]
        x = np.array([[float(blk[i, c]) - float(blk[j, c]) for c in self._cols] for j in others], dtype=float)
        return float(np.mean(self._pairwise_matrix(blk, x, i, others)))

    def potential(self, board: Board, player: Player) -> float:
        probability = min(max(self.predict_player_win_probability(board, player), 1e-6), 1 - 1e-6)
        if self.potential_mode == "logit":
            return float(math.log(probability / (1.0 - probability)))
        return float(probability)


# ---- pairwise training set (analytics/winprob/dataset.py) ------------------------------------------------------------
def resolve_snapshot_path(path: str) -> str:
    """A snapshots file, or the snapshots.parquet / snapshots.csv of a run directory (dataset.py:15-30)."""
    import os

    if os.path.isfile(path):
        return path
    if os.path.isdir(path):
        for name in ("snapshots.parquet", "snapshots.csv"):
            cand = os.path.join(path, name)
            if os.path.exists(cand):
                return cand
    raise FileNotFoundError(f"Could not resolve snapshots dataset from '{path}'. Expected snapshots.parquet or snapshots.csv.")


def load_snapshots_dataframe(path: str):
    import pandas as pd

    resolved = resolve_snapshot_path(path)
    if resolved.lower().endswith(".parquet"):
        return pd.read_parquet(resolved)
    if resolved.lower().endswith(".csv"):
        return pd.read_csv(resolved)
    raise ValueError(f"Unsupported snapshot dataset extension: {resolved}")


phase_bucket_from_occupancy = _phase_bucket_from_occupancy


def build_pairwise_dataset(snapshots_df, *, feature_columns: Sequence[str] = SNAPSHOT_FEATURE_COLUMNS, tie_policy: str = "drop"):
    """dataset.py:57-148: for every (run, game, checkpoint) and every pair of players i < j with different final
    scores, the rows (f_i - f_j, label = score_i > score_j) and the reciprocal (f_j - f_i, 1 - label); tied pairs are
    dropped.  Same rows, order and columns as the reference; the differences are computed per group as one matrix."""
    import pandas as pd

    missing = [c for c in feature_columns if c not in snapshots_df.columns]
    if missing:
        raise ValueError(f"Snapshot dataset missing feature columns: {missing}")
    required = ["run_id", "game_id", "checkpoint_index", "player_id", "agent_name", "final_score", "phase_board_occupancy"]
    missing_required = [c for c in required if c not in snapshots_df.columns]
    if missing_required:
        raise ValueError(f"Snapshot dataset missing required columns: {missing_required}")
    if tie_policy != "drop":
        raise ValueError(f"Unsupported tie_policy '{tie_policy}'. Expected 'drop'.")
    feats = list(feature_columns)
    rows: List[Dict[str, Any]] = []
    ties = 0
    for (run_id, game_id, ck), group in snapshots_df.groupby(["run_id", "game_id", "checkpoint_index"], sort=True):
        players = group.sort_values("player_id")
        if len(players) < 2:
            continue
        occ = float(players["phase_board_occupancy"].mean())
        bucket = _phase_bucket_from_occupancy(occ)
        vals = players[feats].astype(float).to_numpy(dtype=float)
        pids = players["player_id"].astype(int).to_numpy()
        names = players["agent_name"].astype(str).to_numpy()
        scores = players["final_score"].astype(float).to_numpy()
        ii, jj = np.triu_indices(len(players), k=1)                 # pairs in (i, j) lexicographic order
        keep = scores[ii] != scores[jj]
        ties += int((~keep).sum())
        ii, jj = ii[keep], jj[keep]
        diff = vals[ii] - vals[jj]
        label = (scores[ii] > scores[jj]).astype(int)
        for n in range(len(ii)):
            for a, b, d, y in ((ii[n], jj[n], diff[n], label[n]), (jj[n], ii[n], -diff[n], 1 - label[n])):
                row = {"run_id": run_id, "game_id": game_id, "checkpoint_index": int(ck), "phase_board_occupancy": occ,
                       "phase_bucket": bucket, "player_i_id": int(pids[a]), "player_j_id": int(pids[b]),
                       "agent_i": names[a], "agent_j": names[b]}
                row.update(zip(feats, d.tolist()))
                row["label"] = int(y)
                rows.append(row)
    frame = pd.DataFrame(rows)
    meta = {"rows": int(len(frame)), "games": int(frame["game_id"].nunique()) if not frame.empty else 0,
            "checkpoints": int(frame["checkpoint_index"].nunique()) if not frame.empty else 0,
            "ties_dropped": int(ties), "tie_policy": tie_policy}
    return frame, meta


def split_pairwise_by_game(pairwise_df, *, test_size: float, seed: int):
    """dataset.py:151-180: shuffle the sorted game ids with RandomState(seed), first round(n * (1 - test_size)) train."""
    if pairwise_df.empty:
        raise ValueError("Pairwise dataset is empty.")
    games = sorted(pairwise_df["game_id"].unique().tolist())
    np.random.RandomState(seed).shuffle(games)
    cut = max(1, min(int(round(len(games) * (1.0 - test_size))), len(games) - 1))
    train, test = set(games[:cut]), set(games[cut:])
    train_df = pairwise_df[pairwise_df["game_id"].isin(train)].copy()
    test_df = pairwise_df[pairwise_df["game_id"].isin(test)].copy()
    return train_df, test_df, {"train_games": len(train), "test_games": len(test), "train_rows": len(train_df),
                               "test_rows": len(test_df)}


def pairwise_tensors(replay: Mapping[str, Any], game_scores):
    """The same training set straight from device tensors, without a data frame: for replay features [N, K, 4, 34],
    capture mask turn_index >= 0 and final scores [N, 4] (torch tensors on one device) returns (X float64[M, 34],
    y int64[M], game int64[M], checkpoint int64[M]) over all captured snapshots, ordered like build_pairwise_dataset
    (pairs i < j, each followed by its reciprocal; tied pairs dropped)."""
    import torch

    f, ti = replay["features"], replay["turn_index"]
    n, k = ti.shape
    sc = game_scores.to(f.device).to(torch.float64)
    i_idx = torch.tensor([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], device=f.device).view(6, 2)
    a = torch.stack([i_idx[:, 0], i_idx[:, 1]], dim=1).reshape(-1)             # i, j, i, j ... -> rows ij then ji
    b = torch.stack([i_idx[:, 1], i_idx[:, 0]], dim=1).reshape(-1)
    x = f[:, :, a, :] - f[:, :, b, :]                                          # [N, K, 12, 34]
    sa, sb = sc[:, a], sc[:, b]                                                # [N, 12]
    y = (sa > sb).to(torch.int64)[:, None, :].expand(n, k, 12)
    keep = (ti >= 0)[:, :, None] & (sa != sb)[:, None, :]
    g = torch.arange(n, device=f.device)[:, None, None].expand(n, k, 12)
    c = torch.arange(k, device=f.device)[None, :, None].expand(n, k, 12)
    return x[keep], y[keep], g[keep], c[keep]

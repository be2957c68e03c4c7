"""Shared by the oracle (CPU) and kernel (GPU) tests: the searches recorded in tests/golden/mcts.npz by the reference's
own MCTSAgent (tests/golden/generate_golden.py gen_mcts / gen_mcts_deep) and what each one means.

Every entry: root state (+ its incremental frontier sets), and the root's children after the search as
(move id, visits, total_reward), plus [transposition hits, table size].  ``spec(tag)`` returns the search settings."""

import os

import numpy as np

from reinforcementlearning_blokus_b200.state import STATE_DTYPE

REWARD_TOL = 1e-9          # reward sums are float64 sums of integers (exact) or of learned probabilities


def load(golden_dir):
    z = np.load(os.path.join(golden_dir, "mcts.npz"))
    return z, [str(n) for n in z["names"]], int(z["seed"][0])


def root(z, tag):
    return np.frombuffer(z[tag + "/state"].tobytes(), dtype=STATE_DTYPE)[0], z[tag + "/frontier"]


def spec(tag):
    """kind -> dict(iterations, family, and the arguments that differ from the defaults)."""
    kind = tag.split("/")[0]
    return {
        "uct": dict(iterations=200, family="plain"),
        "tt1": dict(iterations=200, family="plain", tt=True),
        "tt2": dict(iterations=200, family="plain", tt=True, second=True),     # same agent, next position, same table
        "heur": dict(iterations=60, family="plain", rollout_policy=1),
        "dummy": dict(iterations=200, family="plain", reward_mode=2, scale=0.5),
        "leaf": dict(iterations=80, family="learned"),
        "bias": dict(iterations=80, family="learned", tt=True, bias_weight=0.25),
        "shape_prob": dict(iterations=40, family="shaped", potential_mode=0),
        "shape_logit": dict(iterations=40, family="shaped", potential_mode=1),
        "gbt": dict(iterations=80, family="learned", model_kind=1),          # pairwise_gbt_phase artifact, leaf evaluation
        "deep_uct": dict(iterations=300, family="plain"),
        "deep_tt": dict(iterations=300, family="plain", tt=True),
        "deep_heur": dict(iterations=150, family="plain", rollout_policy=1, tt=True),
    }[kind]


def compare(z, tag, n_root, moves, visits, reward, scale=1.0):
    """The search's root statistics (by ascending legal-move rank) against the reference's children."""
    got = {int(m): (int(v), float(r) * scale) for m, v, r in zip(moves[:n_root], visits[:n_root], reward[:n_root]) if int(m) != 0xFFFF}
    ref = {int(m): (int(v), float(r)) for m, v, r in zip(z[tag + "/moves"], z[tag + "/visits"], z[tag + "/reward"])}
    assert set(got) == set(ref), (tag, "expanded root moves differ")
    for m, (v, r) in ref.items():
        assert got[m][0] == v, (tag, m, "visits", got[m][0], v)
        assert abs(got[m][1] - r) <= REWARD_TOL * max(1.0, abs(r)), (tag, m, "reward", got[m][1], r)
    # get_best_move(): first child with the most visits in creation order = highest rank among the maxima
    best = max(ref.items(), key=lambda kv: (kv[1][0], kv[0]))[0]
    assert best == int(z[tag + "/best"][0]), tag

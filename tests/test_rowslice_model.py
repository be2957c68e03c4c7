"""CPU model of the row-sliced fit&touch algorithm the CUDA kernels implement.

Interprets the SAME run covers tools/gen_tables.py turns into csrc/orient_gen.cuh, so a wrong
decomposition or a wrong mask rule shows up here, without a GPU, against the pinned oracle."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import gen_tables  # noqa: E402
from oracle import orc  # noqa: E402
from reinforcementlearning_blokus_b200 import pieces  # noqa: E402
from reinforcementlearning_blokus_b200.state import bits_to_int  # noqa: E402

M = (1 << 20) - 1
COVERS = [gen_tables.best_cover(o.cells) for o in pieces.ORIENTATIONS]


def rows_of(mask400):
    return [(mask400 >> (20 * r)) & M for r in range(20)] + [0] * 12


def shd(x, k):  # __shfl_down_sync on zero-padded lanes
    return [x[i + k] if i + k < 32 else x[i] for i in range(32)]


def masks(state, p):
    P = [rows_of(bits_to_int(state["bits"][q])) for q in range(4)]
    own = P[p]
    occ = [P[0][i] | P[1][i] | P[2][i] | P[3][i] for i in range(32)]
    h = [((own[i] << 1) | (own[i] >> 1)) & M for i in range(32)]
    up = lambda x: [0] + x[:-1]
    dn = lambda x: [x[i + 1] if i < 19 else 0 for i in range(32)]
    orth = [h[i] | up(own)[i] | dn(own)[i] for i in range(32)]
    diag = [up(h)[i] | dn(h)[i] for i in range(32)]
    A = [(~(occ[i] | orth[i])) & M if i < 20 else 0 for i in range(32)]
    if (int(state["first_mask"]) >> p) & 1:
        crow = 19 if p >= 2 else 0
        cbit = (1 << 19) if p in (1, 2) else 1
        C = [A[i] & cbit if i == crow else 0 for i in range(32)]
    else:
        C = [A[i] & diag[i] for i in range(32)]
    return A, C


def prim(X, kind, L, r, op):
    Xk = [shd(X, k) for k in range(5)]
    if kind == "H":
        out = Xk[r]
        for j in range(1, L):
            out = [op(out[i], Xk[r][i] >> j) for i in range(32)]
    else:
        out = Xk[r]
        for j in range(1, L):
            out = [op(out[i], Xk[r + j][i]) for i in range(32)]
    return out


def model_moves(state, p):
    A, C = masks(state, p)
    used = int(state["used"][p])
    ids = []
    for o in pieces.ORIENTATIONS:
        if used >> (o.piece_id - 1) & 1:
            continue
        fit = [M] * 32
        touch = [0] * 32
        for kind, L, r, c, _ in COVERS[o.gidx]:
            a = prim(A, kind, L, r, lambda x, y: x & y)
            t = prim(C, kind, L, r, lambda x, y: x | y)
            fit = [fit[i] & (a[i] >> c) for i in range(32)]
            touch = [touch[i] | (t[i] >> c) for i in range(32)]
        for row in range(32):
            Lr = fit[row] & touch[row]
            assert row < 20 or Lr == 0
            for col in range(20):
                if Lr >> col & 1:
                    ids.append(o.gidx * 400 + row * 20 + col)
    return np.array(ids, dtype=np.uint16)


def test_model_matches_reference_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "random_states.npz"))
    states = z["states"].view(orc.STATE_DTYPE).reshape(-1)
    offs, ids = z["offsets"], z["move_ids"]
    for i in range(0, len(states), 3):
        for p in range(4):
            want = ids[offs[4 * i + p]:offs[4 * i + p + 1]]
            got = model_moves(states[i], p)
            assert np.array_equal(got, want), (i, p)


def test_covers_are_exact():
    for o, cov in zip(pieces.ORIENTATIONS, COVERS):
        covered = set()
        for _, _, _, _, cells in cov:
            covered |= set(cells)
        assert covered == set(o.cells)

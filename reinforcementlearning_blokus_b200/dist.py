"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).

Two things shard:
  * independent games -- rank r plays a contiguous block of GLOBAL game indices; the Philox stream of a game
    depends on its global index only, so results do not depend on the number of ranks; no data-path
    collective (scores can be gathered afterwards, 16 B per game);
  * root-parallel MCTS -- every rank searches its own trees from the same root with rank-distinct tree
    indices, then ONE all-reduce(sum) of the root visit / reward vectors (int64, <= 2048 entries each) and
    every rank picks the same move.  The reference has no such step (it is single-process); the merge rule is
    the standard root-parallel one: most summed visits, lowest move id on ties.  On GPUs the all-reduce is replaced by
    ``PeerRootMerge``: one kernel per rank that sums its trees and adds the sums into every rank's window over NVLink peer
    memory (bk_root_merge_p2p) -- same integer sums, no NCCL launch on the move's critical path.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first global index, count) of rank's contiguous block; blocks differ by at most one game."""
    base, extra = divmod(int(n_total), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def sharded_playout(n_total: int, seed: int, game_offset: int = 0, gather_scores: bool = False, **kw) -> Dict[str, torch.Tensor]:
    """Each rank plays its block of the n_total games [game_offset, game_offset+n_total).

    Returns the local playout result plus ``first_game`` / ``count``; with gather_scores every rank also gets
    ``all_game_scores`` int32[n_total, 4] in global game order (all_gather of equal-sized padded blocks)."""
    from . import device as dev

    rank, ws = world()
    start, count = shard_range(n_total, rank, ws)
    res = dev.playout(count, seed=seed, game_offset=game_offset + start, **kw)
    res = dict(res)
    res["first_game"], res["count"] = start, count
    if gather_scores:
        res["all_game_scores"] = gather_rows(res["game_scores"], n_total)
    return res


def sharded_selfplay_mcts(n_total: int, iterations: int, seed: int = 0, game_offset: int = 0, gather_scores: bool = False,
                          **kw) -> Dict[str, torch.Tensor]:
    """MCTS self-play (batched.selfplay_mcts) with the n_total games sharded over the ranks; game g keeps its random
    streams whatever the world size.  No collective on the data path; gather_scores adds all_game_scores on every rank."""
    from . import batched

    rank, ws = world()
    start, count = shard_range(n_total, rank, ws)
    res = dict(batched.selfplay_mcts(count, iterations, seed=seed, game_offset=game_offset + start, **kw))
    res["first_game"], res["count"] = start, count
    if gather_scores:
        res["all_game_scores"] = gather_rows(res["game_scores"], n_total)
    return res


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks (sizes from shard_range) in rank order on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    width = -(-int(n_total) // ws)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r in range(ws):
        _, c = shard_range(n_total, r, ws)
        out.append(parts[r][:c])
    return torch.cat(out, dim=0)


def merge_root_stats(visits: torch.Tensor, reward: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sum root statistics over ranks: the single exchange step of a root-parallel search.

    `visits` / `reward` are this rank's per-move sums (any integer dtype); both come back as int64 sums over all
    ranks.  Integer sums make the result independent of reduction order and of the number of ranks."""
    v = visits.to(torch.int64).clone()
    w = reward.to(torch.int64).clone()
    _, ws = world()
    if ws > 1:
        packed = torch.stack([v, w])            # one collective instead of two
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        v, w = packed[0], packed[1]
    return v, w


def pick_move(visits: torch.Tensor, n_root: int) -> int:
    """Rank of the move with the most summed visits; lowest move id wins ties (same on every rank)."""
    return int(torch.argmax(visits[:n_root]).item())     # argmax returns the first maximal index


def root_parallel_search(root_state: torch.Tensor, sims_total: int, trees_per_rank: int, seed: int,
                         max_rollout_moves: int = 50, exploration: float = 1.414, search_index: int = 0,
                         group=None, peer: Optional["PeerRootMerge"] = None) -> Dict[str, object]:
    """sims_total simulations split evenly over world_size * trees_per_rank trees; one exchange (``peer`` = a
    PeerRootMerge: one kernel over NVLink peer memory; None: all-reduce); every rank returns the same
    {"rank_of_best", "visits", "reward", "n_root", "sims"}."""
    from .mcts import DeviceUCT

    rank, ws = world()
    trees = ws * trees_per_rank
    iters = max(1, int(sims_total) // trees)
    # global tree index = search_index * trees + rank * trees_per_rank + k : disjoint Philox streams
    eng = DeviceUCT(iters, n_trees=trees_per_rank, max_rollout_moves=max_rollout_moves, exploration=exploration,
                    seed=seed, tree_offset=search_index * trees + rank * trees_per_rank)
    res = eng.search(root_state, iters)
    n_root = int(res["counts"][0].item())
    if peer is not None:
        v, w = peer.merge(res["visits"], res["reward"], n_root)
    else:
        v, w = merge_root_stats(res["visits"].sum(dim=0), res["reward"].sum(dim=0), group=group)
    return {"rank_of_best": pick_move(v, n_root), "visits": v[:n_root], "reward": w[:n_root], "n_root": n_root,
            "sims": iters * trees}


class PeerRootMerge:
    """The exchange step of a root-parallel search as ONE kernel per rank over NVLink peer memory (include/blokus_b200.h,
    bk_root_merge_p2p): every rank owns a window (cudaMalloc + CUDA IPC handle), maps the windows of the others, and per
    merge adds the sums over its own trees into all windows with 64-bit atomics, signals, waits for the other ranks inside
    the kernel and reads its own window.  The handles travel once, at construction, through ``all_gather_object``.
    Results are the integer sums ``merge_root_stats`` returns.  One process per GPU; every rank must call ``merge`` the same
    number of times (the epoch is part of the protocol)."""

    def __init__(self, capacity: int = 2048, group=None, timeout_ms: int = 20000):
        from . import _lib
        from . import device as dev

        self.rank, self.ws = world()
        if self.ws > 16:
            raise ValueError("PeerRootMerge supports up to 16 ranks")
        self.capacity, self.timeout_ms, self.epoch, self.group = int(capacity), int(timeout_ms), 0, group
        self.lib = dev._lib_for(None)
        self._check = _lib.check
        own, handle = C.c_void_p(), C.create_string_buffer(64)
        self._check(self.lib.bk_p2p_window_create(self.capacity, C.byref(own), handle))
        self.own = own.value
        handles = [None] * self.ws
        if self.ws > 1:
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.windows = (C.c_void_p * self.ws)()
        self._opened = []
        for r in range(self.ws):
            if r == self.rank:
                self.windows[r] = self.own
            else:
                peer = C.c_void_p()
                self._check(self.lib.bk_p2p_window_open(handles[r], C.byref(peer)))
                self.windows[r] = peer.value
                self._opened.append(peer.value)
        self.merged = torch.zeros((2, self.capacity), dtype=torch.int64, device="cuda")
        self.status = torch.zeros(1, dtype=torch.int32, device="cuda")
        if self.ws > 1:
            dist.barrier(group=group)          # nobody pushes before every window is mapped everywhere

    def merge(self, visits: torch.Tensor, reward: torch.Tensor, n: int, check: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """visits [n_trees, stride] int32, reward [n_trees, stride] int64 (a search's root outputs), n root moves ->
        (visits, reward) int64[n] summed over all trees of all ranks.  More than 32 trees are first added up by torch (the
        kernel is one CTA).  ``check`` reads the status word back (one 4-byte copy); a caller that synchronises anyway can
        pass False and test ``self.status`` itself."""
        from . import device as dev

        if visits.dtype not in (torch.int32, torch.int64) or reward.dtype != torch.int64 or visits.shape != reward.shape or visits.dim() != 2:
            raise ValueError("PeerRootMerge.merge: visits int32/int64 [trees, stride] and reward int64 [trees, stride] expected")
        n = int(n)
        if not 0 <= n <= min(self.capacity, visits.shape[1]):
            raise ValueError("PeerRootMerge.merge: n out of range")
        if visits.shape[0] > 32:
            visits, reward = visits.sum(dim=0, dtype=torch.int64, keepdim=True), reward.sum(dim=0, keepdim=True)
        self._check(self.lib.bk_root_merge_p2p(visits.data_ptr(), int(visits.dtype == torch.int64), reward.data_ptr(), int(visits.shape[0]),
                                               int(visits.stride(0)), n, self.capacity, self.rank, self.ws, self.windows, self.epoch,
                                               self.timeout_ms, self.merged.data_ptr(), self.status.data_ptr(), dev._stream()))
        self.epoch += 1
        out = self.merged.view(-1)[: 2 * n].view(2, n)
        if check and int(self.status.item()) != 0:
            raise RuntimeError("PeerRootMerge: the other ranks did not arrive within %d ms" % self.timeout_ms)
        return out[0], out[1]

    def close(self, collective: bool = True) -> None:
        """Unmap the peers' windows, then free the own one.  ``collective`` (every rank calls close): a barrier in between,
        so that no window is freed while another process still has it mapped."""
        torch.cuda.synchronize()
        for p in self._opened:
            self.lib.bk_p2p_window_close(p)
        self._opened = []
        if self.own:
            if collective and self.ws > 1 and dist.is_initialized():
                dist.barrier(group=self.group)
            self.lib.bk_p2p_window_destroy(self.own)
            self.own = None

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:
            pass


class RootParallelWide:
    """STRONG scaling of one move: a fixed simulation budget split over `total_trees` root-parallel WIDE trees
    (mcts.WideUCT: each tree is itself spread over its GPU), `total_trees / world_size` of them on every rank, one
    exchange of the root statistics -- ``merge="p2p"``: PeerRootMerge (one kernel over NVLink peer memory, the default on
    GPUs), ``merge="nccl"``: torch.distributed all-reduce.  Tree t always has global index search_index * total_trees + t,
    so the merged result is the same for every world size that divides total_trees."""

    def __init__(self, sims_total: int, total_trees: int, seed: int, width: int = 0, group=None, merge: str = "p2p", **wide_kwargs):
        from .mcts import WideUCT

        self.rank, self.ws = world()
        if total_trees % self.ws:
            raise ValueError("total_trees must be a multiple of the world size")
        if merge not in ("p2p", "nccl"):
            raise ValueError("merge must be 'p2p' or 'nccl'")
        self.group = group
        self.total_trees = int(total_trees)
        self.trees_per_rank = self.total_trees // self.ws
        self.per_tree = max(1, int(sims_total) // self.total_trees)
        self.width = int(width) if width else max(64, min(1024, self.per_tree // 4))
        self.eng = WideUCT(self.per_tree, width=self.width, n_trees=self.trees_per_rank, seed=seed,
                           tree_offset=self.rank * self.trees_per_rank, **wide_kwargs)
        self.peer = PeerRootMerge(capacity=int(self.eng.visits.shape[1]), group=group) if merge == "p2p" else None

    def merge(self, res: Dict[str, torch.Tensor], n_root: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.peer is not None:
            return self.peer.merge(res["visits"], res["reward"], n_root)
        v, w = merge_root_stats(res["visits"].sum(dim=0), res["reward"].sum(dim=0), group=self.group)
        return v[:n_root], w[:n_root]

    def search(self, root_state: torch.Tensor, search_index: int = 0) -> Dict[str, object]:
        self.eng.p.tree_offset = int(search_index) * self.total_trees + self.rank * self.trees_per_rank
        res = self.eng.search(root_state, self.per_tree)
        n_root = int(res["counts"][0].item())
        v, w = self.merge(res, n_root)
        return {"rank_of_best": pick_move(v, n_root), "visits": v, "reward": w, "n_root": n_root,
                "sims": self.per_tree * self.total_trees}

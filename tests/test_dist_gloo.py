"""Host-side multi-rank logic on CPU: world_size 2 over gloo (127.0.0.1).  The kernels themselves are
covered by the -m gpu tests; here the sharding arithmetic and the root-statistics exchange are checked."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reinforcementlearning_blokus_b200 import dist as bdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. game sharding: blocks tile [0, n) exactly, global order is preserved by the gather
        n_total = 1001
        start, count = bdist.shard_range(n_total, rank, world)
        local = torch.arange(start, start + count, dtype=torch.int32).unsqueeze(1).repeat(1, 4)
        allrows = bdist.gather_rows(local, n_total)
        ok1 = torch.equal(allrows[:, 0], torch.arange(n_total, dtype=torch.int32)) and allrows.shape == (n_total, 4)
        # 2. root-parallel merge: integer sums, identical on every rank, identical move pick
        g = torch.Generator().manual_seed(100 + rank)
        visits = torch.randint(0, 50, (300,), generator=g, dtype=torch.int32)
        reward = torch.randint(-500, 500, (300,), generator=g, dtype=torch.int64)
        v, w = bdist.merge_root_stats(visits, reward)
        exp_v = torch.zeros(300, dtype=torch.int64)
        exp_w = torch.zeros(300, dtype=torch.int64)
        for r in range(world):
            gg = torch.Generator().manual_seed(100 + r)
            exp_v += torch.randint(0, 50, (300,), generator=gg, dtype=torch.int32).to(torch.int64)
            exp_w += torch.randint(-500, 500, (300,), generator=gg, dtype=torch.int64)
        ok2 = torch.equal(v, exp_v) and torch.equal(w, exp_w)
        # ties: lowest move id
        tie = torch.zeros(10, dtype=torch.int64)
        tie[3] = tie[7] = 5
        ok3 = bdist.pick_move(tie, 10) == 3 and bdist.pick_move(v, 300) == int(torch.argmax(exp_v))
        q.put((rank, ok1, ok2, ok3, bdist.world()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, ok2, ok3, w in sorted(results):
        assert ok1 and ok2 and ok3 and w == (rank, 2)


def test_shard_range_tiles_exactly():
    for n in (0, 1, 7, 65536, 1000003):
        for ws in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(ws):
                s, c = bdist.shard_range(n, r, ws)
                assert s == nxt and c in (n // ws, n // ws + 1)
                nxt += c
            assert nxt == n


def test_single_process_paths():
    assert bdist.world() == (0, 1)
    v, w = bdist.merge_root_stats(torch.tensor([1, 2, 3], dtype=torch.int32), torch.tensor([4, 5, 6]))
    assert v.dtype == torch.int64 and v.tolist() == [1, 2, 3] and w.tolist() == [4, 5, 6]
    x = torch.arange(12).reshape(3, 4)
    assert bdist.gather_rows(x, 3) is x

"""CPU (gloo, world_size 2): host logic of the keyword-sharded multi-GPU path -- shard bounds,
score gather and distributed top-k give exactly the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from enhance_cb_whisper_b200 import parallel


def topk_ref(scores, k, ids=None, id_offset=0):
    """CPU stand-in with the contract of ops.topk: top-k per column, ties -> lower id first."""
    n, U = scores.shape
    if ids is None:
        ids = (torch.arange(n, dtype=torch.int32) + id_offset)[:, None].expand(n, U)
    vals = torch.empty(k, U)
    out_ids = torch.empty(k, U, dtype=torch.int32)
    for u in range(U):
        order = sorted(range(n), key=lambda c: (-float(scores[c, u]), int(ids[c, u])))[:k]
        vals[:, u] = scores[order, u]
        out_ids[:, u] = ids[order, u]
    return vals, out_ids


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 100, 1001):
        for w in (1, 2, 3, 8):
            b = parallel.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, U, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        full = torch.randn(K, U, generator=g)
        if K > 3:
            full[3] = full[1]  # exact ties across shard boundaries
        full[K - 1] = full[0]
        lo, hi = parallel.shard_range(K, world, rank)
        local = full[lo:hi].clone()
        gathered = parallel.gather_scores(local, K)
        v, i = parallel.distributed_topk(local, k, K, topk_ref)
        ev, ei = topk_ref(full, min(k, K))
        q.put((rank, bool(torch.equal(gathered, full)), bool(torch.equal(v, ev)), bool(torch.equal(i, ei))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K,U,k", [(11, 3, 4), (2, 2, 5), (64, 5, 10)])
def test_gather_and_topk_world2(K, U, k):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, K, U, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g_ok, v_ok, i_ok in res:
        assert g_ok, f"rank {rank}: gathered scores differ"
        assert v_ok and i_ok, f"rank {rank}: distributed top-k differs from single-process top-k"


def test_length_balanced_shards_partition_and_balance():
    g = torch.Generator().manual_seed(3)
    for K, w in ((0, 2), (1, 2), (7, 3), (1000, 8), (1001, 8), (64, 2)):
        lens = torch.randint(1, 151, (K,), generator=g)
        if K > 10:
            lens, _ = torch.sort(lens)  # a vocabulary ordered by length: the worst case for a contiguous split
        shards = parallel.length_balanced_shards(lens, w)
        assert len(shards) == w
        allids = torch.cat(shards)
        assert torch.equal(torch.sort(allids).values, torch.arange(K))  # a partition of range(K)
        for s in shards:
            assert torch.equal(s, torch.sort(s).values)  # ascending global ids: the tie-break order is kept
        sizes = [int(s.numel()) for s in shards]
        assert max(sizes) - min(sizes) <= 1
        tot = [int(lens[s].sum()) for s in shards]
        if K:
            assert max(tot) - min(tot) <= int(lens.max())
        if K >= 1000:  # the contiguous split of the same sorted vocabulary is far worse
            cont = [int(lens[lo:hi].sum()) for lo, hi in parallel.shard_bounds(K, w)]
            assert max(cont) - min(cont) > 10 * (max(tot) - min(tot) + 1)


def _worker_indexed(rank, world, port, K, U, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        full = torch.randn(K, U, generator=g)
        if K > 5:
            full[5] = full[2]  # exact ties between keywords that land on different ranks
        full[K - 1] = full[0]
        lens = torch.randint(1, 61, (K,), generator=g)
        shards = parallel.length_balanced_shards(lens, world)
        local = full[shards[rank]].clone()
        gathered = parallel.gather_scores_indexed(local, shards)
        v, i = parallel.distributed_topk_indexed(local, k, shards, topk_ref)
        ev, ei = topk_ref(full, min(k, K))
        q.put((rank, bool(torch.equal(gathered, full)), bool(torch.equal(v, ev)), bool(torch.equal(i, ei))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K,U,k", [(11, 3, 4), (3, 2, 5), (64, 5, 10)])
def test_length_balanced_gather_and_topk_world2(K, U, k):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_indexed, args=(r, 2, port, K, U, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g_ok, v_ok, i_ok in res:
        assert g_ok, f"rank {rank}: gathered scores differ"
        assert v_ok and i_ok, f"rank {rank}: length-balanced distributed top-k differs from single-process top-k"

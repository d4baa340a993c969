"""Keyword-sharded multi-GPU scoring: one process per GPU, keywords partitioned
contiguously across ranks, utterances replicated; no data-path collective until
the final score exchange (every (keyword, utterance) pair is independent --
SURVEY.md section 8e).  ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in
the CPU unit tests of the host logic) is plumbing only.

The reference has no distributed code at all; the exchange implemented here
serves its two consumers of scores: thresholded detections / PR curves over the
full [K, U] matrix (model.py:783-813) and recall@k via top-k over the keyword
axis (model.py:523).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced partition of range(n): the first n % world shards get one extra item."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    return shard_bounds(n, world)[rank]


def gather_scores(local: torch.Tensor, K: int, group=None) -> torch.Tensor:
    """local [K_r, U] scores of this rank's keyword shard -> full [K, U] on every rank
    (uneven shards are padded to the largest shard for the all_gather, then trimmed)."""
    world = dist.get_world_size(group)
    bounds = shard_bounds(K, world)
    kmax = max(hi - lo for lo, hi in bounds)
    U = local.shape[1]
    pad = local.new_zeros((kmax, U))
    pad[: local.shape[0]] = local
    buf = local.new_empty((world, kmax, U))
    dist.all_gather_into_tensor(buf.view(world * kmax, U), pad.contiguous(), group=group)
    return torch.cat([buf[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)], dim=0)


def distributed_topk(local: torch.Tensor, k: int, K: int, topk_fn: Callable, group=None):
    """Per-utterance top-k over the sharded keyword axis.

    local [K_r, U]; ``topk_fn(scores [n,U], k, ids [n,U] | None, id_offset) -> (vals [k,U], ids int32 [k,U])``
    must break ties by the lower global keyword id (``ops.topk`` on GPUs).  Local top-k, all_gather of the
    world*k candidates, merge.  Returns (vals [k,U], global ids [k,U]) identical on every rank and identical
    to a single-device top-k of the gathered matrix.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(K, world, rank)
    assert local.shape[0] == hi - lo, "local shard does not match shard_range"
    kk = min(k, K)
    U = local.shape[1]
    if local.shape[0] > 0:
        v, i = topk_fn(local.contiguous(), min(kk, local.shape[0]), None, lo)
    else:
        v, i = local.new_empty((0, U)), torch.empty((0, U), dtype=torch.int32, device=local.device)
    # pad every rank's candidate list to kk rows (-inf / id -1 never win)
    vp = local.new_full((kk, U), float("-inf"))
    ip = torch.full((kk, U), -1, dtype=torch.int32, device=local.device)
    vp[: v.shape[0]] = v
    ip[: i.shape[0]] = i
    vall = local.new_empty((world * kk, U))
    iall = torch.empty((world * kk, U), dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(vall, vp.contiguous(), group=group)
    dist.all_gather_into_tensor(iall, ip.contiguous(), group=group)
    # -1 ids would win ties against real ids at -inf only; map them past every real id
    iall = torch.where(iall < 0, torch.full_like(iall, 2 ** 31 - 1), iall)
    return topk_fn(vall, kk, iall, 0)


# ---- length-aware sharding (ragged keyword banks) ---------------------------------------------------------
# With the keyword length table carried into the fused kernel (kws_sim_stem_ragged) a keyword costs roughly its number of
# valid frames, so a contiguous split of a vocabulary that is ordered by anything correlated with length (alphabetical
# lists are) leaves ranks unevenly loaded.  SURVEY.md section 8e: sort / bucket keywords by length before sharding.
def length_balanced_shards(lengths: torch.Tensor, world: int) -> List[torch.Tensor]:
    """Partition range(K) into ``world`` shards of (almost) equal size AND equal total length: keywords are sorted
    by length (longest first, stable) and dealt to the ranks in snake order (0..w-1, w-1..0, ...); each shard is
    returned as ascending global keyword ids (int64), so that the local order is still the global order and the
    lower-id tie-break of ``ops.topk`` is preserved.  Shard sizes differ by at most one, total lengths by at most
    the longest keyword."""
    K = int(lengths.numel())
    order = torch.argsort(lengths.reshape(-1).to(torch.int64).cpu(), descending=True, stable=True)
    pos = torch.arange(K)
    rnd, col = pos // world, pos % world
    owner = torch.where(rnd % 2 == 0, col, world - 1 - col)
    return [torch.sort(order[owner == r]).values for r in range(world)]


def gather_scores_indexed(local: torch.Tensor, shards: List[torch.Tensor], group=None) -> torch.Tensor:
    """local [len(shards[rank]), U] scores of this rank's (non-contiguous) shard -> full [K, U] in global keyword
    order on every rank."""
    world = dist.get_world_size(group)
    K = sum(int(s.numel()) for s in shards)
    kmax = max(int(s.numel()) for s in shards)
    U = local.shape[1]
    pad = local.new_zeros((kmax, U))
    pad[: local.shape[0]] = local
    buf = local.new_empty((world * kmax, U))
    dist.all_gather_into_tensor(buf, pad.contiguous(), group=group)
    out = local.new_empty((K, U))
    for r, s in enumerate(shards):
        out[s.to(local.device)] = buf[r * kmax: r * kmax + int(s.numel())]
    return out


def distributed_topk_indexed(local: torch.Tensor, k: int, shards: List[torch.Tensor], topk_fn: Callable, group=None):
    """``distributed_topk`` for non-contiguous shards: the global keyword id of local row i is shards[rank][i]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    K = sum(int(s.numel()) for s in shards)
    kk = min(k, K)
    U = local.shape[1]
    my = shards[rank].to(local.device, torch.int32)
    assert local.shape[0] == my.numel(), "local shard does not match its id list"
    vp = local.new_full((kk, U), float("-inf"))
    ip = torch.full((kk, U), 2 ** 31 - 1, dtype=torch.int32, device=local.device)
    if local.shape[0] > 0:
        ids = my.view(-1, 1).expand(-1, U).contiguous()
        v, i = topk_fn(local.contiguous(), min(kk, local.shape[0]), ids, 0)
        vp[: v.shape[0]] = v
        ip[: i.shape[0]] = torch.where(i < 0, torch.full_like(i, 2 ** 31 - 1), i)
    vall = local.new_empty((world * kk, U))
    iall = torch.empty((world * kk, U), dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(vall, vp.contiguous(), group=group)
    dist.all_gather_into_tensor(iall, ip.contiguous(), group=group)
    return topk_fn(vall, kk, iall, 0)

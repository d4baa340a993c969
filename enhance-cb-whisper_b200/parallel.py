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

"""Corpus sharding across GPUs: one process per GPU, contiguous row ranges, replicated queries.

Every rank searches its own shard (local exact top-k with global row numbers); one ``all_gather`` of the
``[nq, k]`` (score, index) blocks over NCCL / NVLink and a G-way merge kernel (K3) give every rank the
global top-k.  Exact fp64 scores are comparable across shards, so the merge is exact.  The reference has
no multi-GPU path for this stage; this is the design of section 8e.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total, world_size, rank):
    """Contiguous, balanced row range ``[lo, hi)`` of ``rank`` (first ``n_total % world_size`` ranks get one more)."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_topk(scores, idx, group=None):
    """all_gather of local ``[nq, k]`` results -> ``[G, nq, k]`` on every rank (works on gloo and nccl)."""
    world = dist.get_world_size(group)
    s_all = [torch.empty_like(scores) for _ in range(world)]
    i_all = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(s_all, scores.contiguous(), group=group)
    dist.all_gather(i_all, idx.contiguous(), group=group)
    return torch.stack(s_all), torch.stack(i_all)


def merge_reference(scores, idx, k):
    """Host/torch statement of the merge rule (score desc, index asc) used to test the K3 kernel and the
    gloo path: ``scores/idx [G, nq, kk]`` -> ``[nq, k]``."""
    g, nq, kk = scores.shape
    s = scores.permute(1, 0, 2).reshape(nq, g * kk)
    i = idx.permute(1, 0, 2).reshape(nq, g * kk)
    big = torch.iinfo(torch.int64).max
    i_key = torch.where(i < 0, torch.full_like(i, big), i)
    order = torch.argsort(i_key, dim=1, stable=True)
    s, i = torch.gather(s, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(s, dim=1, descending=True, stable=True)
    return torch.gather(s, 1, order)[:, :k], torch.gather(i, 1, order)[:, :k]


def sharded_search(store, queries, k, group=None, **kw):
    """Search this rank's shard, all-gather, merge on the device.  Returns global ``(scores, idx)``."""
    from .engine import merge_topk
    s, i = store.search(queries, k, **kw)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return s, i
    s_all, i_all = gather_topk(s, i, group)
    return merge_topk(s_all, i_all, k)

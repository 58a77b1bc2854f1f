"""Row-sharded search across the GPUs of one box: one process per GPU, torch.distributed (NCCL).

The reference search is single-GPU (it bypasses DataParallel, utils/similarity.py:80-85); its
running top-k (utils/similarity.py:18-35) is a chunk-wise merge, which is what makes the bank
shardable: every rank searches its contiguous row shard, then ONE all-gather of the per-rank
[Q, k] (score, global index) candidates over NVLink is followed by a device merge on every rank.
Bank shards never move.  The first-batch normalisation statistics (utils/similarity.py:98-100)
are computed once and broadcast so every rank normalises identically.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world, align=1):
    """Contiguous item range [lo, hi) of `rank`; shard sizes differ by at most `align` items and
    every boundary is a multiple of `align` (e.g. the synthetic generator's chunk size)."""
    units = (n_items + align - 1) // align
    base, rem = divmod(units, world)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(lo_u * align, n_items), min(hi_u * align, n_items)


def broadcast_norm(bank, src=0, group=None):
    """Make rank `src`'s first-batch statistics the statistics of every shard."""
    mu, sigma = bank.norm() if dist.get_rank(group) == src else (
        torch.empty(bank.D, device=bank.device), torch.empty(bank.D, device=bank.device))
    dist.broadcast(mu, src, group=group)
    dist.broadcast(sigma, src, group=group)
    bank.set_norm(mu, sigma)
    return mu, sigma


def gather_candidates(scores, idx, group=None):
    """All-gather per-rank candidates [Q, k] -> [world, Q, k] (scores f32, idx i64)."""
    world = dist.get_world_size(group)
    g_s = torch.empty((world,) + tuple(scores.shape), device=scores.device, dtype=scores.dtype)
    g_i = torch.empty((world,) + tuple(idx.shape), device=idx.device, dtype=idx.dtype)
    if scores.is_cuda:
        dist.all_gather_into_tensor(g_s, scores.contiguous(), group=group)
        dist.all_gather_into_tensor(g_i, idx.contiguous(), group=group)
    else:   # gloo (CPU tests of the host logic)
        dist.all_gather(list(g_s.unbind(0)), scores.contiguous(), group=group)
        dist.all_gather(list(g_i.unbind(0)), idx.contiguous(), group=group)
    return g_s, g_i


def sharded_search(local_search, k, metric, merge=None, group=None):
    """local_search() -> (scores [Q, k], idx [Q, k] with GLOBAL indices) on this rank's shard.
    Returns the global top-k on every rank.  `merge` defaults to the CUDA merge kernel."""
    scores, idx = local_search()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return scores, idx
    g_s, g_i = gather_candidates(scores, idx, group)
    if merge is None:
        from .engine import merge_candidates as merge
    return merge(g_s, g_i, k, metric)


class ShardedBank:
    """A Bank holding rows [row_lo, row_hi) of a global bank of n_total items."""

    def __init__(self, bank, row_lo, n_total, group=None):
        self.bank, self.row_lo, self.n_total, self.group = bank, int(row_lo), int(n_total), group

    def search(self, t, w=None, k=100, metric="cosine", combine="min", n_top_sims=None, path="auto"):
        return sharded_search(
            lambda: self.bank.search(t, w, k, metric, combine, n_top_sims, path, idx_offset=self.row_lo),
            k, metric, group=self.group)

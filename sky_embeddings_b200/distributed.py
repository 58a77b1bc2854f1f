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


class CandidateExchange:
    """The exchange step of a sharded search with ONE collective: a rank's [Q, k] indices (i64) and scores (f32)
    live in one preallocated device buffer, so a single NCCL all-gather moves both, and the merge kernel reads
    the gathered per-rank blocks in place (sky_merge_candidates_strided) -- no packing or copy kernels."""

    def __init__(self, Q, k, device, group=None, world=None):
        self.Q, self.k, self.group = int(Q), int(k), group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        n = self.Q * self.k
        self.units = n + (n + 1) // 2                      # int64 units per rank: idx | scores (two f32 per unit)
        self.local = torch.empty(self.units, dtype=torch.int64, device=device)
        self.gathered = torch.empty(self.world * self.units, dtype=torch.int64, device=device)
        self.idx = self.local[:n].view(self.Q, self.k)
        self.scores = self.local[n:].view(torch.float32)[:n].view(self.Q, self.k)
        self.out_scores = torch.empty((self.Q, self.k), dtype=torch.float32, device=device)
        self.out_idx = torch.empty((self.Q, self.k), dtype=torch.int64, device=device)

    def merge(self, metric):
        """All-gather the local candidates (already written into .scores / .idx) and merge; returns the global
        (scores, idx), identical on every rank."""
        if self.world == 1:
            return self.scores, self.idx
        dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
        return self.merge_gathered(metric)

    def merge_gathered(self, metric):
        """Merge the per-rank blocks already sitting in .gathered (rank r at [r * units, (r + 1) * units))."""
        import ctypes as C
        from . import _lib as L
        from .engine import _stream
        n = self.Q * self.k
        base = self.gathered.data_ptr()
        L.check(L.load().sky_merge_candidates_strided(
            C.c_void_p(base + n * 8), C.c_void_p(base), self.world, self.Q, self.k, self.units * 2, self.units,
            self.k, L.METRICS["cosine" if metric == "cosine" else "MSE"], C.c_void_p(self.out_scores.data_ptr()),
            C.c_void_p(self.out_idx.data_ptr()), self.gathered.device.index, _stream(self.gathered.device)))
        return self.out_scores, self.out_idx


class PeerExchange:
    """The exchange step over PEER MEMORY instead of a collective: every rank's candidate block is pushed into a
    buffer of every peer with plain NVLink stores and a per-query flag, and the merge kernel waits for its flags
    (csrc/exchange.cu).  torch.distributed is used once, at construction, to trade the 64-byte cudaIpc handles.
    Same interface as CandidateExchange: write the local candidates into .scores / .idx, call merge()."""

    def __init__(self, Q, k, device, group=None):
        import ctypes as C
        from . import _lib as L
        self.Q, self.k, self.group = int(Q), int(k), group
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PeerExchange needs CUDA devices with peer access: there is no CPU fallback")
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        lib = L.load()
        self._h = C.c_void_p()
        L.check(lib.sky_exchange_create(C.byref(self._h), self.device.index or 0, self.rank, self.world, self.Q, self.k))
        if self.world > 1:
            nb = lib.sky_exchange_handle_bytes()
            mine = (C.c_ubyte * nb)()
            L.check(lib.sky_exchange_handle(self._h, mine))
            local = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
            every = torch.empty(self.world * nb, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(every, local, group=group)
            blob = bytes(every.cpu().tolist())
            L.check(lib.sky_exchange_open(self._h, blob))
            dist.barrier(group=group)                  # nobody pushes before every peer has mapped every buffer
        self.scores = torch.empty((self.Q, self.k), dtype=torch.float32, device=self.device)
        self.idx = torch.empty((self.Q, self.k), dtype=torch.int64, device=self.device)
        self.out_scores = torch.empty((self.Q, self.k), dtype=torch.float32, device=self.device)
        self.out_idx = torch.empty((self.Q, self.k), dtype=torch.int64, device=self.device)

    def merge(self, metric):
        import ctypes as C
        from . import _lib as L
        from .engine import _stream
        if self.world == 1:
            return self.scores, self.idx
        L.check(L.load().sky_exchange_merge(
            self._h, C.c_void_p(self.scores.data_ptr()), C.c_void_p(self.idx.data_ptr()), self.Q, self.k, self.k,
            L.METRICS["cosine" if metric == "cosine" else "MSE"], C.c_void_p(self.out_scores.data_ptr()),
            C.c_void_p(self.out_idx.data_ptr()), _stream(self.device)))
        return self.out_scores, self.out_idx

    def search(self, bank, t, w=None, metric="cosine", combine="min", n_top_sims=None, path="auto", idx_offset=0):
        """The fused route: `bank`'s shard search delivers its result straight into every peer (the shard merge kernel
        writes over NVLink), then the flag-waiting merge -- sky_search_sharded.  Returns the global (scores, idx)."""
        if self.world == 1:
            return bank.search(t, w, self.k, metric, combine, n_top_sims, path, idx_offset=idx_offset,
                               out_scores=self.out_scores, out_idx=self.out_idx)
        return bank.search_sharded(self._h, t, w, self.k, metric, combine, n_top_sims, path, idx_offset=idx_offset,
                                   out_scores=self.out_scores, out_idx=self.out_idx)

    def close(self):
        if getattr(self, "_h", None):
            from . import _lib as L
            torch.cuda.synchronize(self.device)
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)         # peers may still be reading what this rank maps
            L.load().sky_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                from . import _lib as L
                L.load().sky_exchange_destroy(self._h)
                self._h = None
        except Exception:
            pass


def make_exchange(Q, k, device, group=None, kind="peer"):
    """kind: 'peer' (NVLink stores + flags, csrc/exchange.cu) or 'nccl' (one all-gather + merge kernel)."""
    if kind == "peer":
        return PeerExchange(Q, k, device, group)
    if kind == "nccl":
        return CandidateExchange(Q, k, device, group)
    raise ValueError(f"unknown exchange kind {kind!r}")


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

    def __init__(self, bank, row_lo, n_total, group=None, exchange="peer"):
        self.bank, self.row_lo, self.n_total, self.group = bank, int(row_lo), int(n_total), group
        self.exchange = exchange
        self._xchg = None

    def search(self, t, w=None, k=100, metric="cosine", combine="min", n_top_sims=None, path="auto"):
        if dist.is_initialized() and dist.get_world_size(self.group) > 1 and t.is_cuda:
            Q = 1 if t.dim() == 1 else t.shape[0]
            if self._xchg is None or (self._xchg.Q, self._xchg.k) != (Q, k):
                if self._xchg is not None and hasattr(self._xchg, "close"):
                    self._xchg.close()
                self._xchg = make_exchange(Q, k, self.bank.device, self.group, self.exchange)
            x = self._xchg
            if isinstance(x, PeerExchange):
                s, i = x.search(self.bank, t, w, metric, combine, n_top_sims, path, idx_offset=self.row_lo)
            else:
                self.bank.search(t, w, k, metric, combine, n_top_sims, path, idx_offset=self.row_lo,
                                 out_scores=x.scores, out_idx=x.idx)
                s, i = x.merge(metric)
            return s.clone(), i.clone()
        return sharded_search(
            lambda: self.bank.search(t, w, k, metric, combine, n_top_sims, path, idx_offset=self.row_lo),
            k, metric, group=self.group)

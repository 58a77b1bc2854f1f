"""CPU port of the reference search loop, for TIMING the reference's algorithm on host cores.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and `--impl reference` legs, and tests).
The reference is Python and /root/reference does not exist on the GPU box, so its CPU path is
restated here with the same torch operations in the same order as
/root/reference/utils/similarity.py, including its waste, because that is what a user of the
reference pays for:
  * per batch: normalise with the first batch's statistics            (:98-102)
  * per batch: recompute the target mean / inverse-variance weights   (:245 -> :134-147)
  * metric with full [B, L, D] temporaries                            (:163-170, :188-192, :208-212)
  * running top-k by cat + argsort + gather of scores, ra_dec AND the payload rows (:18-35)
torch decides the thread count (all host cores unless told otherwise).
tests/test_ref_port.py checks this port against the golden fixtures produced by the reference.
"""
from __future__ import annotations

import torch


def _token_view(latent, n_extra, cls_token, max_pool):
    if cls_token:
        return latent[:, :1]
    kept = latent[:, n_extra:]
    return kept.max(dim=1, keepdim=True).values if max_pool else kept


def _group_stats(group):
    rows = group.reshape(-1, group.shape[-1])
    centre = rows.mean(dim=0)
    inv_var = rows.std(dim=0).pow(2).reciprocal()
    return centre, inv_var / inv_var.sum()


def _cos(t, x, w, eps=1e-6):
    num = (w * t * x).sum(-1)
    den = (w * t.pow(2)).sum(-1).sqrt() * (w * x.pow(2)).sum(-1).sqrt() + eps
    return num / den


def _mse(t, x, w):
    return ((t - x).pow(2) * w / w.sum()).mean(-1)


def _mae(t, x, w):
    return ((t - x).abs() * w / w.sum()).mean(-1)


_METRIC = {"cosine": _cos, "MSE": _mse, "MAE": _mae}
_COMBINE = {"mean": lambda s: s.mean(1), "min": lambda s: s.min(1).values, "max": lambda s: s.max(1).values}


def batch_scores(target, batch, metric, combine, use_weights, n_top_sims=None):
    t, w = _group_stats(target)
    if not use_weights:
        w = torch.ones_like(w)
    s = _METRIC[metric](t, batch, w)
    if n_top_sims is not None:
        s = s.topk(n_top_sims, dim=1, largest=(metric == "cosine")).values
    return _COMBINE[combine](s)


def running_topk(keep, new, n_save, descending):
    """keep/new: tuples (scores, ra_dec, payload); returns the merged best n_save."""
    merged = [torch.cat((a, b), 0) for a, b in zip(keep, new)]
    order = torch.argsort(merged[0], descending=descending)[:n_save]
    return tuple(m[order] for m in merged)


def search_loop(target_latent, bank_latent, batch_size, n_save, metric="cosine", combine="min",
                use_weights=True, max_pool=False, cls_token=False, n_extra=1, payload=None):
    """One reference-style pass over a pre-encoded bank [N, tokens, D] (the ViT forward is out of
    scope and is not timed).  payload: rows gathered alongside the scores like the reference's image
    tensor (defaults to the latents themselves).  Returns (scores[n_save], idx[n_save])."""
    descending = metric == "cosine"
    target = _token_view(target_latent, n_extra, cls_token, max_pool)
    n = bank_latent.shape[0]
    payload = bank_latent if payload is None else payload
    best = (torch.full((n_save,), float("-inf") if descending else float("inf")),
            torch.zeros((n_save, 2)),
            torch.zeros((n_save, *payload.shape[1:]), dtype=payload.dtype))
    mu = sd = None
    with torch.no_grad():
        for start in range(0, n, batch_size):
            stop = min(n, start + batch_size)
            batch = _token_view(bank_latent[start:stop], n_extra, cls_token, max_pool)
            if mu is None:
                mu = batch.mean(dim=(0, 1))
                sd = batch.std(dim=(0, 1), unbiased=True)
                target = (target - mu) / (sd + 1e-8)
            batch = (batch - mu) / (sd + 1e-8)
            scores = batch_scores(target, batch, metric, combine, use_weights)
            ra = torch.zeros((stop - start, 2))
            ra[:, 0] = torch.arange(start, stop, dtype=torch.float32)
            best = running_topk(best, (scores, ra, payload[start:stop]), n_save, descending)
    return best[0], best[1][:, 0].to(torch.int64)


def multi_query_loop(queries_t, queries_w, bank_z, batch_size, n_save, metric):
    """The multi-query workload of the bench, the way a reference user would have to run it: the
    reference handles ONE aggregated query per pass, so Q queries are Q passes.
    queries_t/w: [Q, D] already-prepared (t, w); bank_z: normalised [N, 1, D].  Each pass keeps the
    reference's per-batch structure (metric temporaries + cat/argsort top-k)."""
    descending = metric == "cosine"
    out = []
    with torch.no_grad():
        for q in range(queries_t.shape[0]):
            t = queries_t[q]
            w = queries_w[q] if queries_w is not None else torch.ones_like(t)
            best_s = torch.full((n_save,), float("-inf") if descending else float("inf"))
            best_i = torch.zeros((n_save,), dtype=torch.int64)
            for start in range(0, bank_z.shape[0], batch_size):
                batch = bank_z[start:start + batch_size]
                s = _COMBINE["min"](_METRIC[metric](t, batch, w))
                ids = torch.arange(start, start + batch.shape[0])
                best_s, best_i = running_topk((best_s, best_i), (s, ids), n_save, descending)
            out.append((best_s, best_i))
    return out

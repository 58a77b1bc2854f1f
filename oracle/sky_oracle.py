"""CPU oracle for the sky_embeddings similarity-search path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sky_embeddings_b200/`` may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs use it, and only as the checker / the reported CPU baseline.

It is a numpy restatement (float64 by default) of the reference's algorithm in
``/root/reference/utils/similarity.py``; each function cites the lines it follows.
Parity pin: ``oracle/make_golden.py`` executes the *unmodified reference functions*
(imported by path in the build container) on seeded inputs and stores their outputs
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures, so the oracle is pinned to the reference itself (the
reference ships no tests or golden vectors of its own).

Conventions
-----------
* ``z`` is a *normalised* bank of shape [N, L, D] (N items, L tokens kept per item).
* A query is the pair ``(t, w)`` of target mean vector and per-feature weights, both [D]
  (or stacked [Q, D] for the multi-query generalisation, whose oracle is the
  single-query reference applied once per query).
* Selection order: best first; NaN scores rank as the largest value (first for
  cosine, last for MSE/MAE) exactly like ``torch.argsort``; exact ties are broken
  by the lower bank index (the reference's unstable sort leaves tie order
  unspecified, so comparisons against it are tie-aware).
"""
from __future__ import annotations

import numpy as np

METRICS = ("cosine", "MSE", "MAE")
COMBINES = ("mean", "min", "max")
NORM_EPS = 1e-8       # utils/similarity.py:101-102
COSINE_EPS = 1e-6     # utils/similarity.py:149 (eps default)
PIXEL_EPS = 1e-5      # SURVEY.md section 8(d) pixel-space definition


def largest_is_best(metric: str) -> bool:
    """utils/similarity.py:233-236 and :20-29 -- cosine sorts descending, MSE/MAE ascending."""
    if metric == "cosine":
        return True
    if metric in ("MSE", "MAE"):
        return False
    # The reference falls through to an UnboundLocalError (:250-259); the oracle is explicit.
    raise ValueError(f"unknown metric {metric!r}; the reference accepts 'cosine', 'MSE', 'MAE'")


def token_select(latent, num_extra_tokens=1, cls_token=False, max_pool=False):
    """Token slicing / max-pool applied to target and bank latents alike.

    utils/similarity.py:55-63 (targets) and :87-95 (bank).  latent: [B, 1+P(+1), D].
    """
    latent = np.asarray(latent)
    if cls_token:
        return latent[:, :1]
    latent = latent[:, num_extra_tokens:]
    if max_pool:
        latent = latent.max(axis=1, keepdims=True)
    return latent


def first_batch_stats(bank_tokens, norm_rows, dtype=np.float64):
    """Mean and *unbiased* std over (items, tokens) of the first ``norm_rows`` items.

    utils/similarity.py:98-100.  bank_tokens: [N, L, D] after token_select.
    A single sample gives NaN std, like torch.
    """
    first = np.asarray(bank_tokens[:norm_rows], dtype=dtype)
    flat = first.reshape(-1, first.shape[-1])
    mu = flat.mean(axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        if flat.shape[0] > 1:
            sigma = flat.std(axis=0, ddof=1)
        else:
            sigma = np.full(flat.shape[1], np.nan, dtype=dtype)
    return mu, sigma


def normalise(x, mu, sigma, dtype=np.float64):
    """utils/similarity.py:101-102 -- (x - mean) / (std + 1e-8)."""
    x = np.asarray(x, dtype=dtype)
    return (x - mu.astype(dtype)) / (sigma.astype(dtype) + dtype(NORM_EPS))


def target_features(target, use_weights=True, dtype=np.float64):
    """Query vector and per-feature weights from a (normalised) target group.

    utils/similarity.py:134-147 (+ :246-247 for use_weights=False).
    target: [T, L_t, D] -> t[D], w[D].
    """
    target = np.asarray(target, dtype=dtype)
    flat = target.reshape(-1, target.shape[-1])
    t = flat.mean(axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        if flat.shape[0] > 1:
            std = flat.std(axis=0, ddof=1)
        else:
            std = np.full(flat.shape[1], np.nan, dtype=dtype)
        w = 1.0 / std ** 2
        w = w / w.sum()
    if not use_weights:
        w = np.ones_like(w)
    return t, w


def token_scores(t, w, z, metric):
    """Per-token score of every bank token against one query.

    cosine: utils/similarity.py:163-170   MSE: :188-192   MAE: :208-212.
    t, w: [D]; z: [N, L, D] -> [N, L].
    """
    largest_is_best(metric)
    z = np.asarray(z)
    dtype = z.dtype
    t = np.asarray(t, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    D = z.shape[-1]
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        if metric == "cosine":
            dot = (w * t * z).sum(axis=-1)
            mag_t = np.sqrt((w * t ** 2).sum(axis=-1))
            mag_z = np.sqrt((w * z ** 2).sum(axis=-1))
            return dot / (mag_t * mag_z + dtype.type(COSINE_EPS))
        if metric == "MSE":
            return (((t - z) ** 2) * w / w.sum()).sum(axis=-1) / D
        return (np.abs(t - z) * w / w.sum()).sum(axis=-1) / D


def combine_scores(s, metric, combine="mean", n_top_sims=None):
    """Optional best-n patch selection then mean/min/max over patches.

    utils/similarity.py:257-267.  s: [N, L] -> [N].  NaN propagates like torch
    (torch.min/max/mean return NaN if any element is NaN).
    """
    s = np.asarray(s)
    if n_top_sims is not None:
        if n_top_sims > s.shape[1]:
            raise RuntimeError("selected index k out of range")  # torch.topk's error
        srt = np.sort(s, axis=1)          # NaN last == "largest", as torch.topk treats it
        s = srt[:, ::-1][:, :n_top_sims] if largest_is_best(metric) else srt[:, :n_top_sims]
    with np.errstate(invalid="ignore"):
        if combine == "mean":
            return s.mean(axis=1)
        if combine == "min":
            return s.min(axis=1)
        if combine == "max":
            return s.max(axis=1)
    # The reference silently returns the un-combined [N, L] tensor here (:262-268).
    raise ValueError(f"unknown combine {combine!r}")


def item_scores(t, w, z, metric, combine="mean", n_top_sims=None):
    """compute_similarity after determine_target_features (utils/similarity.py:250-268)."""
    return combine_scores(token_scores(t, w, z, metric), metric, combine, n_top_sims)


def compute_similarity(target, test, metric="MAE", combine="mean", use_weights=True,
                       n_top_sims=None, dtype=np.float64):
    """Full restatement of utils/similarity.py:214-268 (n_central_patches excluded:
    it raises NameError in the reference)."""
    t, w = target_features(target, use_weights, dtype)
    return item_scores(t, w, np.asarray(test, dtype=dtype), metric, combine, n_top_sims)


def order_best_first(scores, metric):
    """Permutation that sorts scores best-first: NaN counts as the largest value
    (utils/similarity.py:24,29 via torch.argsort), ties by lower index."""
    scores = np.asarray(scores, dtype=np.float64)
    nan = np.isnan(scores)
    idx = np.arange(scores.shape[0])
    if largest_is_best(metric):
        key = np.where(nan, np.inf, scores)
        # NaN must outrank +inf: sort by (not nan, -key, idx)
        return np.lexsort((idx, -key, ~nan))
    key = np.where(nan, np.inf, scores)
    return np.lexsort((idx, key, nan))


def topk(scores, k, metric):
    """Global top-k, best first, padded like the reference when N < k
    (utils/similarity.py:65-66: -inf for cosine, +inf otherwise; index -1)."""
    scores = np.asarray(scores)
    order = order_best_first(scores, metric)[:k]
    out_s = scores[order].astype(np.float64)
    out_i = order.astype(np.int64)
    if order.shape[0] < k:
        pad = k - order.shape[0]
        fill = -np.inf if largest_is_best(metric) else np.inf
        out_s = np.concatenate([out_s, np.full(pad, fill)])
        out_i = np.concatenate([out_i, np.full(pad, -1, dtype=np.int64)])
    return out_s, out_i


def search(t, w, z, k, metric="cosine", combine="min", n_top_sims=None):
    """Multi-query exact top-k over a normalised bank.

    t, w: [Q, D] (w None = unweighted, i.e. ones); z: [N, L, D].
    Returns scores [Q, k] float64, idx [Q, k] int64 (best first).
    The oracle of the multi-query engine is the single-query reference once per query.
    """
    t = np.atleast_2d(np.asarray(t))
    Q = t.shape[0]
    z = np.asarray(z)
    if w is None:
        w = np.ones_like(t)
    w = np.atleast_2d(np.asarray(w))
    out_s = np.empty((Q, k), dtype=np.float64)
    out_i = np.empty((Q, k), dtype=np.int64)
    for q in range(Q):
        s = item_scores(t[q].astype(z.dtype), w[q].astype(z.dtype), z, metric, combine, n_top_sims)
        out_s[q], out_i[q] = topk(s, k, metric)
    return out_s, out_i


def simsearch(target_latent, bank_latent, norm_rows, k, metric="cosine", combine="min",
              use_weights=True, max_pool=False, cls_token=False, num_extra_tokens=1,
              dtype=np.float64):
    """End-to-end restatement of mae_simsearch (utils/similarity.py:37-132) over a
    pre-encoded bank: token select, first-batch normalisation, target features,
    scoring, global top-k.  Returns (scores[k], idx[k], t[D], w[D], mu[D], sigma[D]).
    """
    tgt = token_select(target_latent, num_extra_tokens, cls_token, max_pool)
    bank = token_select(bank_latent, num_extra_tokens, cls_token, max_pool)
    mu, sigma = first_batch_stats(bank, norm_rows, dtype)
    tgt_n = normalise(tgt, mu, sigma, dtype)
    z = normalise(bank, mu, sigma, dtype)
    t, w = target_features(tgt_n, use_weights, dtype)
    s = item_scores(t, w, z, metric, combine)
    sc, ix = topk(s, k, metric)
    return sc, ix, t, w, mu, sigma


def merge_topk(score_lists, index_lists, k, metric):
    """Merge per-shard / per-batch candidate lists into one top-k
    (the decomposition utils/similarity.py:18-35 performs batch by batch)."""
    s = np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in score_lists])
    i = np.concatenate([np.asarray(a, dtype=np.int64).ravel() for a in index_lists])
    keep = i >= 0
    s, i = s[keep], i[keep]
    nan = np.isnan(s)
    key = np.where(nan, np.inf, s)
    if largest_is_best(metric):
        order = np.lexsort((i, -key, ~nan))[:k]
    else:
        order = np.lexsort((i, key, nan))[:k]
    out_s, out_i = s[order], i[order]
    if order.shape[0] < k:
        pad = k - order.shape[0]
        fill = -np.inf if largest_is_best(metric) else np.inf
        out_s = np.concatenate([out_s, np.full(pad, fill)])
        out_i = np.concatenate([out_i, np.full(pad, -1, dtype=np.int64)])
    return out_s, out_i


def pixel_masked_mse(q, x, qmask=None, dtype=np.float64):
    """Pixel-space masked MSE (BASELINE config 5).

    No such search exists in the reference; it is defined so that the reference's
    weighted_MSE (utils/similarity.py:174-192) with w = mask is the oracle up to the
    eps term, with the NaN handling of forward_loss (utils/mim_vit.py:482-486,509-519):
      valid = ~isnan(q) & ~isnan(x);  m = valid * qmask;  NaN -> 0 after masking
      score = sum(m (q - x)^2) / (sum(m) + 1e-5)       (lower is better)
    q: [C,H,W] or [D]; x: [N, ...] -> [N].
    """
    q = np.asarray(q, dtype=dtype).reshape(-1)
    x = np.asarray(x, dtype=dtype).reshape(x.shape[0], -1)
    m = (~np.isnan(q))[None, :] & ~np.isnan(x)
    if qmask is not None:
        m = m & (np.asarray(qmask).reshape(-1) != 0)[None, :]
    q0 = np.nan_to_num(q, nan=0.0)
    x0 = np.nan_to_num(x, nan=0.0)
    d = (q0[None, :] - x0) ** 2 * m
    return d.sum(axis=1) / (m.sum(axis=1) + dtype(PIXEL_EPS))


# ----------------------------------------------------------------------------------------
# comparison helpers (tie-aware), shared by the CPU and GPU parity tests
# ----------------------------------------------------------------------------------------

def score_close(a, b, rel, scale=None):
    """|a-b| <= rel * max(|b|, scale) elementwise, NaN==NaN, inf==inf (SURVEY.md section 7:
    tolerance must be scale-relative because cosine -> 0 makes pure relative error blow up)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if scale is None:
        fin = np.isfinite(b)
        scale = float(np.max(np.abs(b[fin]))) if fin.any() else 1.0
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= rel * np.maximum(np.abs(b), scale)
    return ok | both_nan | same_inf


def check_topk_parity(got_scores, got_idx, ref_scores, ref_idx, rel, all_scores=None, scale=None):
    """Tie-aware top-k comparison (SURVEY.md section 8(c) iv).

    * scores must match position-wise within ``rel`` (scale-relative);
    * indices must be identical wherever the reference's adjacent score gaps exceed the
      tolerance; inside a tie group (gap <= tol) any permutation is accepted, and at the
      k-boundary an index may be swapped for another bank row whose score ties with the
      k-th (checked through ``all_scores`` when given).
    Returns (ok, message).
    """
    got_scores = np.asarray(got_scores, dtype=np.float64)
    ref_scores = np.asarray(ref_scores, dtype=np.float64)
    got_idx = np.asarray(got_idx, dtype=np.int64)
    ref_idx = np.asarray(ref_idx, dtype=np.int64)
    if got_scores.shape != ref_scores.shape:
        return False, f"shape {got_scores.shape} vs {ref_scores.shape}"
    fin = np.isfinite(ref_scores)
    if scale is None:
        scale = float(np.max(np.abs(ref_scores[fin]))) if fin.any() else 1.0
    ok = score_close(got_scores, ref_scores, rel, scale)
    if not ok.all():
        j = int(np.argmin(ok))
        return False, f"score mismatch at rank {j}: got {got_scores[j]!r} ref {ref_scores[j]!r}"
    k = ref_scores.shape[0]
    tol = 2.0 * rel * scale
    # group consecutive ranks whose reference scores are within tol of each other
    groups, start = [], 0
    for j in range(1, k + 1):
        boundary = j == k
        if not boundary:
            a, b = ref_scores[j - 1], ref_scores[j]
            same = (np.isnan(a) and np.isnan(b)) or (np.isinf(a) and np.isinf(b) and a == b) \
                or (np.isfinite(a) and np.isfinite(b) and abs(a - b) <= tol)
            boundary = not same
        if boundary:
            groups.append((start, j))
            start = j
    for (lo, hi) in groups:
        g, r = set(got_idx[lo:hi].tolist()), set(ref_idx[lo:hi].tolist())
        if g == r:
            continue
        last_group = hi == k
        if last_group and all_scores is not None:
            # boundary ties: extra indices are fine if their true score ties with the group
            extra = [i for i in g - r if i >= 0]
            lo_s, hi_s = np.nanmin(ref_scores[lo:hi]) - tol, np.nanmax(ref_scores[lo:hi]) + tol
            vals = np.asarray(all_scores, dtype=np.float64)[extra]
            if len(extra) == len(g - r) and np.all((vals >= lo_s) & (vals <= hi_s)):
                continue
        return False, f"index mismatch in ranks [{lo},{hi}): got {sorted(g)} ref {sorted(r)}"
    return True, "ok"

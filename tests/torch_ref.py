"""fp32 torch restatement of the reference scores over a STORED bank, for full-size parity gates.

Checker only (tests/ and bench.py's pre-timing gate): at BASELINE sizes (1M ... 50M rows) the numpy oracle is too
slow, so the reference formulas (utils/similarity.py:163-170 cosine, :188-192 MSE; SURVEY.md section 8(d) for the
pixel-space masked MSE) are evaluated chunk by chunk with plain torch fp32 ops on the rows the bank actually stores
(bank.download), with the UNROUNDED fp32 queries, and the running top-k is kept with torch.topk.
"""
from __future__ import annotations

import torch


def _scores(z, t, w, metric, D):
    """z [n, D] f32 stored rows, t [q, D], w [q, D] or None -> [q, n] reference scores in fp32."""
    if metric == "cosine":
        if w is None:
            num = t @ z.T
            den = t.pow(2).sum(1).sqrt()[:, None] * z.pow(2).sum(1).sqrt()[None, :] + 1e-6
        else:
            num = (w * t) @ z.T
            den = (w * t * t).sum(1).sqrt()[:, None] * (w @ z.pow(2).T).sqrt() + 1e-6
        return num / den
    # direct form, one query at a time (no cancellation): sum w (t - z)^2 / (D sum w)
    out = torch.empty((t.shape[0], z.shape[0]), device=z.device, dtype=torch.float32)
    for q in range(t.shape[0]):
        d = (z - t[q]).pow(2)
        if w is None:
            out[q] = d.sum(1) / (D * D)
        else:
            out[q] = (d * w[q]).sum(1) / (D * w[q].sum())
    return out


def fp32_topk(bank, t, w, metric, k, idx_offset=0, step=1 << 17):
    """Reference top-k (scores [q, k], idx [q, k]) of queries t (and weights w) over every stored row of `bank`."""
    dev = bank.device
    t = t.to(dev, torch.float32)
    w = None if w is None else w.to(dev, torch.float32)
    q, D, n = t.shape[0], bank.D, bank.n_items
    largest = metric == "cosine"
    kk = min(k, n)
    best_s = torch.full((q, kk), float("-inf") if largest else float("inf"), device=dev)
    best_i = torch.full((q, kk), -1, dtype=torch.int64, device=dev)
    for s0 in range(0, n, step):
        z = bank.download(s0, min(step, n - s0))[:, 0]
        s = _scores(z, t, w, metric, D)
        cs = torch.cat([best_s, s], 1)
        ci = torch.cat([best_i, torch.arange(s0, s0 + z.shape[0], device=dev).expand(q, -1) + idx_offset], 1)
        best_s, o = cs.topk(kk, dim=1, largest=largest)
        best_i = ci.gather(1, o)
    return best_s, best_i


def pixel_scores(x, q, qmask=None):
    """Masked MSE of one query cutout q [D] against rows x [n, D] (NaN = missing), SURVEY.md section 8(d)."""
    m = ~torch.isnan(x) & ~torch.isnan(q)[None, :]
    if qmask is not None:
        m &= (qmask != 0)[None, :]
    d = torch.where(m, x - torch.nan_to_num(q)[None, :], torch.zeros((), device=x.device))
    return d.pow(2).sum(1) / (m.sum(1).to(torch.float32) + 1e-5)


def check_topk(got_s, got_i, ref_s, ref_i, rel, largest, scale=None):
    """Tie-aware comparison of one query's top-k against the reference's.  Returns (ok, message, max_err) with
    max_err = the largest |got - ref| / max(|ref|, scale) over the ranks.
      * scores rank by rank within rel * max(|ref|, scale)   (scale defaults to max |ref|: cosine -> 0 makes a pure
        relative error meaningless);
      * best first;
      * the index sets agree except inside tie groups: an index missing on either side must sit within the tolerance
        of the k-th score."""
    got_s, ref_s = got_s.double().cpu(), ref_s.double().cpu()
    got_i, ref_i = got_i.cpu(), ref_i.cpu()
    k = ref_s.shape[0]
    if got_s.shape[0] != k:
        return False, f"length {got_s.shape[0]} != {k}", float("inf")
    sc = float(ref_s.abs().max()) if scale is None else float(scale)
    tol = rel * torch.maximum(ref_s.abs(), torch.tensor(sc, dtype=torch.float64))
    err = (got_s - ref_s).abs()
    max_err = float((err / torch.maximum(ref_s.abs(), torch.tensor(sc, dtype=torch.float64))).max())
    if bool((err > tol).any()):
        j = int((err - tol).argmax())
        return False, f"score rank {j}: got {float(got_s[j])!r} ref {float(ref_s[j])!r} (tol {float(tol[j]):.3g})", max_err
    d = got_s[1:] - got_s[:-1]
    if bool(((d > 0) if largest else (d < 0)).any()):
        return False, "scores are not best-first", max_err
    kth, tk = float(ref_s[-1]), float(tol[-1])
    gs, rs = set(got_i.tolist()), set(ref_i.tolist())
    for i in gs - rs:          # returned but not in the reference list: must tie with the k-th score
        s = float(got_s[got_i.tolist().index(i)])
        if abs(s - kth) > 2 * tk:
            return False, f"index {i} (score {s!r}) is not in the reference top-{k} (k-th {kth!r})", max_err
    for i in rs - gs:
        s = float(ref_s[ref_i.tolist().index(i)])
        if abs(s - kth) > 2 * tk:
            return False, f"reference index {i} (score {s!r}) is missing (k-th {kth!r})", max_err
    return True, "ok", max_err


def merge_sorted_lists(scores, idx, k, largest):
    """Bit-exact model of the device merge: candidates [R, q, k_in] -> top-k by (score best first, index ascending)."""
    R, q, kin = scores.shape
    s = scores.permute(1, 0, 2).reshape(q, R * kin)
    i = idx.permute(1, 0, 2).reshape(q, R * kin)
    o1 = torch.argsort(i, dim=1, stable=True)
    s1, i1 = s.gather(1, o1), i.gather(1, o1)
    o2 = torch.sort(s1, dim=1, descending=largest, stable=True).indices
    return s1.gather(1, o2)[:, :k].contiguous(), i1.gather(1, o2)[:, :k].contiguous()

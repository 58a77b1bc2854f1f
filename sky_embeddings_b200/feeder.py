"""Embedding feeder: encoder output -> device-resident bank, without the host round trip.

The reference extracts embeddings with ``mae_latent`` (/root/reference/utils/eval_fns.py:72-140),
which moves every batch of latents to the CPU (``latents.append(latent.detach().cpu())``, :132) and
concatenates them there; ``mae_simsearch`` (/root/reference/utils/similarity.py:71-102) re-encodes the
whole bank for every target instead.  Here each batch goes from the encoder (still the reference's
PyTorch model -- the ViT is out of scope) straight into the bank: token select / max-pool, the
first-batch normalisation (:98-102), the cast and the tile-major store are one ingest kernel, and the
bank then serves any number of searches.
"""
from __future__ import annotations

import torch

from .engine import Bank, token_mode_of, tokens_kept
from .similarity import get_train_samples, select_tokens


def bank_from_loader(model, dataloader, device, n_batches=None, max_pool=False, cls_token=False,
                     nested_batches=False, bank_dtype="bf16", n_items=None, keep_samples=False, verbose=0):
    """Encode every batch of ``dataloader`` and build the resident bank.

    model / dataloader follow the duck-typed contracts of mae_simsearch (utils/similarity.py:41-52,
    :71-85).  n_items: capacity of the bank (default ``len(dataloader.dataset)``; when unknown the
    latents are gathered on the device first).  Returns ``(bank, ra_decs [N, 2] on device,
    samples [N, C, H, W] on the host or None)``; bank row i is loader item i (``shuffle=False``,
    similarity_search.py:155-156).
    """
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("bank_from_loader needs device='cuda': sky_embeddings_b200 has no CPU fallback")
    model.eval()
    enc = model.module if hasattr(model, "module") else model
    n_extra = enc.num_extra_tokens
    mode = token_mode_of(max_pool, cls_token)
    if n_items is None and not nested_batches and hasattr(dataloader, "dataset"):
        try:
            n_items = len(dataloader.dataset)
        except TypeError:
            n_items = None
    bank, done, pending = None, 0, []
    ra_all, img_all = [], []
    first = None
    with torch.no_grad():
        for i, (samples, _masks, ra_decs) in enumerate(get_train_samples(dataloader, nested_batches)):
            samples = samples.to(device, non_blocking=True)
            ra_decs = ra_decs.to(device, non_blocking=True)
            latent, _, _ = enc.forward_features(samples, ra_dec=ra_decs, reshape_out=False)
            B, tokens, D = latent.shape
            if first is None:
                first = latent            # the first batch defines the normalisation statistics
            if n_items is not None:
                if bank is None:
                    bank = Bank(n_items, tokens_kept(tokens, mode, n_extra), D, bank_dtype, device)
                    bank.fit_norm(first, mode, n_extra)
                bank.upload(latent, done, mode, n_extra)
            else:
                pending.append(latent)
            done += B
            ra_all.append(ra_decs)
            if keep_samples:
                img_all.append(samples.cpu())
            if verbose and (i + 1) % verbose == 0:
                print(f"Encoded {i + 1} batches...", end="\r")
            if n_batches is not None and (i + 1) >= n_batches:
                break
    if done == 0:
        raise ValueError("the loader yielded no batches")
    if bank is None:
        lat = torch.cat(pending)
        bank = Bank(lat.shape[0], tokens_kept(lat.shape[1], mode, n_extra), lat.shape[2], bank_dtype, device)
        bank.fit_norm(first, mode, n_extra)
        bank.upload(lat, 0, mode, n_extra)
    elif done < bank.n_items:
        bank.resize(done)
    bank.finalize()
    return bank, torch.cat(ra_all), (torch.cat(img_all) if keep_samples else None)


def resident_simsearch(bank, target_latent, ra_decs, samples=None, num_extra_tokens=1, n_save=256, metric="cosine",
                       combine="min", use_weights=True, max_pool=False, cls_token=False):
    """mae_simsearch over a resident bank: one search instead of a pass over the loader.
    Returns (best_samples or None, best_idx [n_save] i64, best_ra_decs [n_save, 2], best_scores [n_save]),
    best first -- the reference's 4-tuple with the winners' bank indices in place of their re-encoded
    latents (re-encode ``best_samples`` with the model if those are needed, utils/similarity.py:124-130)."""
    tsel = select_tokens(target_latent.to(bank.device), num_extra_tokens, cls_token, max_pool)
    t, w = bank.query_from_targets(tsel, use_weights)
    scores, idx = bank.search(t, w if use_weights else None, k=n_save, metric=metric, combine=combine)
    scores, idx = scores[0], idx[0]
    ok = idx >= 0
    safe = idx.clamp(min=0)
    best_ra = torch.where(ok[:, None], ra_decs[safe], torch.zeros_like(ra_decs[safe]))
    best_samples = None
    if samples is not None:
        best_samples = samples[safe.cpu()]
        best_samples[~ok.cpu()] = 0
    return best_samples, idx, best_ra, scores

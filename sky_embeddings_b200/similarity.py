"""Drop-in mirror of the reference's search interface, backed by libskysearch.so.

Same function names, argument meaning, return layout and error behaviour as
/root/reference/utils/similarity.py, so that

    from sky_embeddings_b200.similarity import mae_simsearch, compute_similarity

can replace ``from utils.similarity import mae_simsearch, compute_similarity``
(/root/reference/similarity_search.py:14, sky_sim_search.py:14).  All scoring, selection and
merging runs in the CUDA library; torch is used for tensors, the encoder call (which stays the
reference's own PyTorch model) and payload gathers.  There is no CPU path: tensors must live on a
CUDA device, otherwise the calls raise.
"""
from __future__ import annotations

import time

import torch

from . import _lib as L
from .engine import Bank, _ptr, _stream, merge_candidates, token_mode_of, tokens_kept


class UnknownMetricError(UnboundLocalError, ValueError):
    """The reference dies with UnboundLocalError on an unknown metric
    (utils/similarity.py:250-259); this is that error, made explicit."""


def _check_metric(metric):
    if metric not in L.METRICS:
        raise UnknownMetricError(f"unknown metric {metric!r}: expected 'cosine', 'MSE' or 'MAE'")


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: sky_embeddings_b200 has no CPU fallback")


def get_train_samples(dataloader, nested_batches):
    """Flat loaders yield batches; tile loaders yield one sky tile already cut into batches
    (reference utils/similarity.py:4-14)."""
    if not nested_batches:
        yield from ((s, m, r) for s, m, r in dataloader)
        return
    for tile_samples, tile_masks, tile_ra_decs in dataloader:
        yield from zip(tile_samples[0], tile_masks[0], tile_ra_decs[0])


def select_tokens(latent, num_extra_tokens=1, cls_token=False, max_pool=False):
    """Token slicing / max-pool of utils/similarity.py:55-63 and :87-95, done by the ingest kernel."""
    _require_cuda(latent, "latent")
    mode = token_mode_of(max_pool, cls_token)
    n, tokens, D = latent.shape
    tmp = Bank(n, tokens_kept(tokens, mode, num_extra_tokens), D, "fp32", latent.device)
    tmp.upload(latent, 0, mode, num_extra_tokens)
    out = tmp.download()
    tmp.close()
    return out


def determine_target_features(target_latent):
    """(avg_feat, weight_feat) of a target group -- reference utils/similarity.py:134-147."""
    _require_cuda(target_latent, "target_latent")
    D = target_latent.shape[-1]
    x = target_latent.to(torch.float32).reshape(-1, D).contiguous()
    t = torch.empty(D, device=x.device, dtype=torch.float32)
    w = torch.empty_like(t)
    lib = L.load()
    L.check(lib.sky_query_from_targets(None, _ptr(x), x.shape[0], D, 1, _ptr(t), _ptr(w), _stream(x.device)))
    return t, w


def _token_scores(target_feats, test_feats, weights, metric):
    _require_cuda(test_feats, "test_feats")
    shape = test_feats.shape[:-1]
    D = test_feats.shape[-1]
    rows = test_feats.reshape(-1, 1, D)
    bank = Bank.from_latents(rows, None, "all", 0, "fp32", test_feats.device)
    s = bank.score(target_feats, weights, metric, "mean")[0]
    bank.close()
    return s.reshape(shape)


def weighted_cosine_similarity(target_feats, test_feats, weights, eps=1e-6):
    """reference utils/similarity.py:149-172 -> [batch, patches]"""
    if eps != 1e-6:
        raise NotImplementedError("the fused kernel fixes eps at the reference default 1e-6")
    return _token_scores(target_feats, test_feats, weights, "cosine")


def weighted_MSE(target_feats, test_feats, weights):
    """reference utils/similarity.py:174-192 -> [batch, patches]"""
    return _token_scores(target_feats, test_feats, weights, "MSE")


def weighted_MAE(target_feats, test_feats, weights):
    """reference utils/similarity.py:194-212 -> [batch, patches]"""
    return _token_scores(target_feats, test_feats, weights, "MAE")


def compute_similarity(target_latent, test_latent, metric='MAE', combine='mean', use_weights=True,
                       n_central_patches=None, n_top_sims=None):
    """One similarity value per sample of test_latent -- reference utils/similarity.py:214-268.

    target_latent [T, L_t, D], test_latent [B, L, D] -> [B].
    """
    _check_metric(metric)
    if n_central_patches is not None:
        # the reference never imports select_centre (utils/misc.py:99) and fails the same way
        raise NameError("name 'select_centre' is not defined")
    _require_cuda(test_latent, "test_latent")
    target_latent = target_latent.to(test_latent.device)
    bank = Bank.from_latents(test_latent, None, "all", 0, "fp32", test_latent.device)
    t, w = bank.query_from_targets(target_latent, use_weights)
    if combine not in L.COMBINES:
        # the reference silently returns the un-combined [B, L] matrix (:262-268)
        shape = test_latent.shape[:-1]
        bank.close()
        rows = Bank.from_latents(test_latent.reshape(-1, 1, test_latent.shape[-1]), None, "all", 0, "fp32",
                                 test_latent.device)
        out = rows.score(t, w if use_weights else None, metric, "mean")[0].reshape(shape)
        rows.close()
        return out
    out = bank.score(t, w if use_weights else None, metric, combine, n_top_sims)[0]
    bank.close()
    return out


def update_best_scores(samples, ra_decs, similarity_scores, best_samples, best_ra_decs, best_scores,
                       n_save, metric):
    """Running top-n_save merge -- reference utils/similarity.py:18-35 (cat + argsort + slice),
    here one device merge of two candidate lists followed by payload gathers."""
    _require_cuda(similarity_scores, "similarity_scores")
    dev = similarity_scores.device
    metric_key = "cosine" if metric == "cosine" else "MSE"     # every non-cosine metric sorts ascending
    nb, nn = best_scores.shape[0], similarity_scores.shape[0]
    width = max(nb, nn)
    fill = float("-inf") if metric == "cosine" else float("inf")
    sc = torch.full((2, 1, width), fill, device=dev, dtype=torch.float32)
    ix = torch.full((2, 1, width), -1, device=dev, dtype=torch.int64)
    sc[0, 0, :nb] = best_scores
    sc[1, 0, :nn] = similarity_scores
    ix[0, 0, :nb] = torch.arange(nb, device=dev)
    ix[1, 0, :nn] = torch.arange(nb, nb + nn, device=dev)
    k = min(n_save, nb + nn)
    out_s, out_i = merge_candidates(sc, ix, k, metric_key)
    order = out_i[0]
    best_scores = out_s[0]
    best_samples = torch.cat((best_samples, samples), dim=0)[order]
    best_ra_decs = torch.cat((best_ra_decs, ra_decs), dim=0)[order]
    return best_samples, best_ra_decs, best_scores


def mae_simsearch(model, target_latent, dataloader, device, n_batches=None,
                  metric='cosine', combine='min', use_weights=True, max_pool=False,
                  cls_token=False, nested_batches=True, n_save=256, verbose=100, bank_dtype="fp32"):
    """Streaming similarity search -- reference utils/similarity.py:37-132.

    Returns (best_samples [n_save, C, H, W], best_latent [n_save, tokens, D],
             best_ra_decs [n_save, 2], best_scores [n_save]) best first; as in the reference, slots
    beyond the number of bank items keep the +-inf initial score (their payload is zero here,
    uninitialised memory there).  The encoder stays the reference's PyTorch model; each batch of
    latents is token-selected, normalised with the FIRST batch's statistics (:98-102), scored and
    merged on the device by libskysearch.  ``bank_dtype`` is the only extra argument.
    """
    _check_metric(metric)
    if combine not in L.COMBINES:
        raise ValueError(f"unknown combine {combine!r}: expected 'mean', 'min' or 'max'")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("mae_simsearch needs device='cuda': sky_embeddings_b200 has no CPU fallback")
    if not nested_batches:
        if n_batches is None:
            n_batches = len(dataloader)
        print(f'Performing similarity search on {min(len(dataloader), n_batches)} batches...')
    else:
        print(f'Performing similarity search on {len(dataloader)} tiles...')
    model.eval()
    enc = model.module if hasattr(model, 'module') else model
    num_extra_tokens = enc.num_extra_tokens
    mode = token_mode_of(max_pool, cls_token)

    target_latent = target_latent.to(device, non_blocking=True)
    target_sel = select_tokens(target_latent, num_extra_tokens, cls_token, max_pool)

    fill = float('-inf') if metric == 'cosine' else float('inf')
    best_scores = torch.full((n_save,), fill, device=device)
    best_ra_decs = torch.zeros((n_save, 2), device=device)
    best_samples = None
    bank = None
    t = w = None
    time_start = time.time()
    with torch.no_grad():
        for i, (samples, masks, ra_decs) in enumerate(get_train_samples(dataloader, nested_batches)):
            samples = samples.to(device, non_blocking=True)
            ra_decs = ra_decs.to(device, non_blocking=True)
            if i == 0:
                best_samples = torch.zeros((n_save, *samples.shape[1:]), device=device, dtype=samples.dtype)
            test_latent, _, _ = enc.forward_features(samples, ra_dec=ra_decs, reshape_out=False)
            B, tokens, D = test_latent.shape
            if bank is None or B > bank_capacity:
                # the first batch defines capacity and the normalisation statistics
                new_bank = Bank(B, tokens_kept(tokens, mode, num_extra_tokens), D, bank_dtype, device)
                if bank is None:
                    new_bank.fit_norm(test_latent, mode, num_extra_tokens)
                    t, w = new_bank.query_from_targets(target_sel, use_weights)
                else:
                    new_bank.set_norm(*bank.norm())
                    bank.close()
                bank, bank_capacity = new_bank, B
            bank.resize(B)
            bank.upload(test_latent, 0, mode, num_extra_tokens).finalize()
            kb = min(n_save, B)
            sc, ix = bank.search(t, w if use_weights else None, kb, metric, combine)
            # running merge with the best so far (reference update_best_scores, :108-110)
            width = max(n_save, kb)
            cs = torch.full((2, 1, width), fill, device=device)
            ci = torch.full((2, 1, width), -1, device=device, dtype=torch.int64)
            cs[0, 0, :n_save] = best_scores
            ci[0, 0, :n_save] = torch.arange(n_save, device=device)
            cs[1, 0, :kb] = sc[0]
            ci[1, 0, :kb] = ix[0] + n_save
            ms, mi = merge_candidates(cs, ci, n_save, "cosine" if metric == "cosine" else "MSE")
            order = mi[0]
            best_scores = ms[0]
            best_samples = torch.cat((best_samples, samples), dim=0)[order]
            best_ra_decs = torch.cat((best_ra_decs, ra_decs.to(best_ra_decs.dtype)), dim=0)[order]

            if not nested_batches:
                if (i + 1) % verbose == 0:
                    print(f'Processed {i+1}/{n_batches} image batches...', end='\r')
                if (i + 1) >= n_batches:
                    break
            elif (i + 1) % verbose == 0:
                time_per_batch = (time.time() - time_start) / (i + 1)
                print(f'Processed {i+1} image batches ({time_per_batch:0.2f} seconds per batch)...', end='\r')

        if bank is not None:
            bank.close()
        best_latent, _, _ = enc.forward_features(best_samples, ra_dec=best_ra_decs, reshape_out=False)
    return best_samples, best_latent, best_ra_decs, best_scores


def save_results(path, test_ra_decs, test_scores, target_images, target_latent, test_images, test_latent):
    """The reference's .npz result layout (similarity_search.py:178-181, sky_sim_search.py:171-174)."""
    import numpy as np
    np.savez(path,
             test_ra_decs=test_ra_decs.data.cpu().numpy(), test_scores=test_scores.data.cpu().numpy(),
             target_images=target_images.data.cpu().numpy(), target_features=target_latent.data.cpu().numpy(),
             test_images=test_images.data.cpu().numpy(), test_features=test_latent.data.cpu().numpy())

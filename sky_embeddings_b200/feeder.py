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
                       combine="min", use_weights=True, max_pool=False, cls_token=False, model=None):
    """mae_simsearch over a resident bank: one search instead of a pass over the loader.
    Returns (best_samples or None, best_idx [n_save] i64, best_ra_decs [n_save, 2], best_scores [n_save]),
    best first -- the reference's 4-tuple with the winners' bank indices in place of their re-encoded latents.
    With ``model`` (and ``samples``) the winners are re-encoded as the reference does (utils/similarity.py:124-130) and
    the second element is ``best_latent`` [n_save, 1 + P, D]: exactly the reference's return value."""
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
    if model is not None:
        if best_samples is None:
            raise ValueError("resident_simsearch(model=...) re-encodes the winners: pass the bank's `samples` as well")
        enc = model.module if hasattr(model, "module") else model
        best_samples = best_samples.to(bank.device)
        with torch.no_grad():
            best_latent, _, _ = enc.forward_features(best_samples, ra_dec=best_ra, reshape_out=False)
        return best_samples, best_latent, best_ra, scores
    return best_samples, idx, best_ra, scores


# ---------------------------------------------------------------------------------------------------------------
# target side: mae_latent with target augmentation, latents kept on the device
# ---------------------------------------------------------------------------------------------------------------
class _RandomBrightness:
    """img * U(lo, hi) -- reference utils/dataloaders.py:13-24"""

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def __call__(self, img):
        import random
        return img * random.uniform(self.lo, self.hi)


class _RandomNoise:
    """img + randn * U(lo, hi) -- reference utils/dataloaders.py:26-37"""

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def __call__(self, img):
        import random
        return img + torch.randn_like(img) * random.uniform(self.lo, self.hi)


class _RandomChannelNaN:
    """0..max_channels random channels set to NaN -- reference utils/dataloaders.py:39-87 (on a copy: the reference
    writes into its argument, which is always a fresh tensor there because the crop precedes it)."""

    def __init__(self, max_channels):
        self.max_channels = max_channels

    def __call__(self, img):
        import random
        C = img.shape[-3]
        if self.max_channels > C:
            raise ValueError(f"max_channels must be less than or equal to the number of channels in the image. "
                             f"Got {self.max_channels} for an image with {C} channels.")
        out = img.clone()
        if out.dim() == 3:
            for c in random.sample(range(C), random.randint(0, self.max_channels)):
                out[c] = torch.nan
        else:
            for b in range(out.shape[0]):
                for c in random.sample(range(C), random.randint(0, self.max_channels)):
                    out[b, c] = torch.nan
        return out


def get_augmentations(img_size=64, flip=True, crop=True, brightness=0.8, noise=0.01, nan_channels=2):
    """The reference's augmentation pipeline (utils/dataloaders.py:90-106), same transforms and ranges, usable on CUDA
    tensors: random flips, RandomResizedCrop(scale 0.8-1, ratio 0.9-1.1), brightness, noise, NaN channels."""
    from torchvision.transforms import v2
    t = []
    if flip:
        t += [v2.RandomHorizontalFlip(), v2.RandomVerticalFlip()]
    if crop:
        t.append(v2.RandomResizedCrop(size=(img_size, img_size), scale=(0.8, 1.0), ratio=(0.9, 1.1), antialias=True))
    if brightness is not None:
        t.append(_RandomBrightness(brightness, 1 / brightness))
    if noise is not None:
        t.append(_RandomNoise(0.0, noise))
    if nan_channels is not None:
        t.append(_RandomChannelNaN(nan_channels))
    return v2.Compose(t)


def mae_latent(model, dataloader, device, n_batches=None, return_images=False, verbose=1,
               apply_augmentations=False, num_augmentations=16, remove_cls=True, augmentations=None):
    """Mirror of the reference's ``mae_latent`` (utils/eval_fns.py:72-140), same arguments: every sample followed by
    its ``num_augmentations`` augmented copies (:92-108), encoded batch by batch.  Differences, both on purpose:
    the augmentations run on the DEVICE tensors, and latents / images stay on the device (the reference moves every
    batch to the host, :132-134, and ``mae_simsearch`` moves the targets straight back).  ``augmentations`` overrides
    the pipeline (e.g. a deterministic one in tests)."""
    device = torch.device(device)
    if n_batches is None:
        n_batches = len(dataloader)
    if verbose > 0:
        print(f'Encoding {min(len(dataloader), n_batches)} batches...')
    model.eval()
    enc = model.module if hasattr(model, "module") else model
    aug = (augmentations or get_augmentations()) if apply_augmentations else None
    latents, images = [], []
    with torch.no_grad():
        for samples, _masks, ra_decs in dataloader:
            samples = samples.to(device, non_blocking=True)
            ra_decs = ra_decs.to(device, non_blocking=True)
            if aug is not None:
                rows, rds = [], []
                for idx in range(samples.shape[0]):
                    rows.append(samples[idx:idx + 1])
                    rows += [aug(samples[idx]).unsqueeze(0) for _ in range(num_augmentations)]
                    rds += [ra_decs[idx:idx + 1]] * (1 + num_augmentations)
                samples, ra_decs = torch.cat(rows), torch.cat(rds)
            latent, _, _ = enc.forward_features(samples, ra_dec=ra_decs, mask=None, reshape_out=False)
            if getattr(enc, "attn_pool", False):
                remove_cls = False
            if remove_cls:
                latent = latent[:, enc.num_extra_tokens:]
            latents.append(latent.detach())
            if return_images:
                images.append(samples.detach())
            if len(latents) >= n_batches:
                break
    if return_images:
        return torch.cat(latents), torch.cat(images)
    return torch.cat(latents)

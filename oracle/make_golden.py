"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

It imports /root/reference/utils/similarity.py by path (it depends on torch only) and drives
  * mae_simsearch      (utils/similarity.py:37-132) through a latent-cache stub model and a
                       sliceable loader, so the reference code runs unchanged over a
                       pre-encoded bank; bank indices are recovered from ra_dec[:, 0];
  * compute_similarity (:214-268) directly;
  * update_best_scores (:18-35) directly, including NaN / inf / tie inputs.
Inputs come from sky_embeddings_b200.synth (seeded numpy PCG64); the fixtures store the seeds,
an input checksum and the reference's outputs.  The reference cannot travel to the GPU box, the
fixtures do.  Nothing here is imported by the product.
"""
from __future__ import annotations

import contextlib
import hashlib
import importlib.util
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sky_embeddings_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

from oracle.ref_harness import LatentLoader, LatentStub, load_reference, quiet  # noqa: E402


def checksum(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def simsearch_cases(ref):
    """mae_simsearch over small latent banks: every metric x token mode x combine x weights."""
    out = {}
    n, P, D, bs, k = 600, 4, 48, 64, 20
    seed = synth.BASE_SEED
    bank = synth.latents(n, 1 + P, D, seed=seed, stream=11)
    tgt = synth.target_group(bank, [17, 333], copies=7, noise=0.3, seed=seed, stream=12)
    out["meta"] = np.array([n, P, D, bs, k, seed, 11, 12, 7], dtype=np.int64)
    out["anchors"] = np.array([17, 333], dtype=np.int64)
    out["noise"] = np.float32(0.3)
    out["checksum"] = np.array(checksum(bank, tgt))
    names = []
    modes = [("maxpool", dict(max_pool=True, cls_token=False), ["min"]),
             ("cls", dict(max_pool=False, cls_token=True), ["min"]),
             ("patches", dict(max_pool=False, cls_token=False), ["mean", "min", "max"])]
    for metric in ("cosine", "MSE", "MAE"):
        for mname, mkw, combines in modes:
            for combine in combines:
                for uw in (True, False):
                    name = f"{metric}.{mname}.{combine}.{'w' if uw else 'u'}"
                    res = quiet(ref.mae_simsearch, LatentStub(1), torch.from_numpy(tgt),
                                LatentLoader(bank, bs), "cpu", metric=metric, combine=combine,
                                use_weights=uw, nested_batches=False, n_save=k, **mkw)
                    best_samples, best_latent, best_ra, best_scores = res
                    idx = best_ra[:, 0].numpy().astype(np.int64)
                    assert np.array_equal(best_samples.numpy(), bank[idx]), name
                    out[f"scores.{name}"] = best_scores.numpy()
                    out[f"idx.{name}"] = idx
                    names.append(name)
    out["names"] = np.array(names)
    return out


def mim1_shape_case(ref):
    """BASELINE config 1 shape: D=768 (configs/mim_1.ini:26), max-pooled (L=1), cosine,
    use_weights=True, k=10, batch 64 -- on synthetic latents (no ViT weights offline)."""
    out = {}
    n, P, D, bs, k = 2000, 8, 768, 64, 10
    seed = synth.BASE_SEED
    bank = synth.latents(n, 1 + P, D, seed=seed, stream=21)
    tgt = synth.target_group(bank, [5, 1234], copies=65, noise=0.5, seed=seed, stream=22)
    out["meta"] = np.array([n, P, D, bs, k, seed, 21, 22, 65], dtype=np.int64)
    out["anchors"] = np.array([5, 1234], dtype=np.int64)
    out["noise"] = np.float32(0.5)
    out["checksum"] = np.array(checksum(bank, tgt))
    for metric in ("cosine", "MSE"):
        res = quiet(ref.mae_simsearch, LatentStub(1), torch.from_numpy(tgt), LatentLoader(bank, bs),
                    "cpu", metric=metric, combine="min", use_weights=True, max_pool=True,
                    cls_token=False, nested_batches=False, n_save=k)
        out[f"scores.{metric}"] = res[3].numpy()
        out[f"idx.{metric}"] = res[2][:, 0].numpy().astype(np.int64)
    return out


def short_bank_case(ref):
    """N < n_save: the +-inf initial fill leaks into the result (utils/similarity.py:65-66)."""
    out = {}
    n, P, D, bs, k = 12, 2, 16, 8, 20
    seed = synth.BASE_SEED
    bank = synth.latents(n, 1 + P, D, seed=seed, stream=31)
    tgt = synth.target_group(bank, [3], copies=5, noise=0.3, seed=seed, stream=32)
    out["meta"] = np.array([n, P, D, bs, k, seed, 31, 32, 5], dtype=np.int64)
    out["checksum"] = np.array(checksum(bank, tgt))
    for metric in ("cosine", "MSE"):
        res = quiet(ref.mae_simsearch, LatentStub(1), torch.from_numpy(tgt), LatentLoader(bank, bs),
                    "cpu", metric=metric, combine="mean", use_weights=True, max_pool=False,
                    cls_token=False, nested_batches=False, n_save=k)
        out[f"scores.{metric}"] = res[3].numpy()
        # rows >= n are torch.empty garbage in the reference: keep only the valid prefix
        out[f"idx.{metric}"] = res[2][:n, 0].numpy().astype(np.int64)
    return out


def compute_similarity_cases(ref):
    out = {}
    seed = synth.BASE_SEED
    T, L, D, B = 9, 6, 32, 50
    target = synth.latents(T, L, D, seed=seed, stream=41)
    test = synth.latents(B, L, D, seed=seed, stream=42)
    out["meta"] = np.array([T, L, D, B, seed, 41, 42], dtype=np.int64)
    out["checksum"] = np.array(checksum(target, test))
    names = []
    for metric in ("cosine", "MSE", "MAE"):
        for combine in ("mean", "min", "max"):
            for uw in (True, False):
                for nts in (None, 2):
                    name = f"{metric}.{combine}.{'w' if uw else 'u'}.{nts}"
                    s = ref.compute_similarity(torch.from_numpy(target), torch.from_numpy(test),
                                               metric=metric, combine=combine, use_weights=uw,
                                               n_top_sims=nts)
                    out[f"scores.{name}"] = s.numpy()
                    names.append(name)
    out["names"] = np.array(names)
    t, w = ref.determine_target_features(torch.from_numpy(target))
    out["t"] = t.numpy()
    out["w"] = w.numpy()
    return out


def update_best_cases(ref):
    """update_best_scores with NaN / inf / ties: fixes the NaN-is-largest ordering."""
    out = {}
    best = torch.tensor([0.9, 0.5, float("-inf"), float("-inf")])
    new = torch.tensor([0.5, float("nan"), 0.7, float("inf"), -1.0, 0.5])
    pay_b = torch.arange(4, dtype=torch.float32).view(4, 1)
    pay_n = torch.arange(4, 10, dtype=torch.float32).view(6, 1)
    ra_b = torch.stack([torch.arange(4.), torch.zeros(4)], 1)
    ra_n = torch.stack([torch.arange(4., 10.), torch.zeros(6)], 1)
    out["best"] = best.numpy()
    out["new"] = new.numpy()
    for metric in ("cosine", "MSE"):
        b = best if metric == "cosine" else -best
        s, ra, sc = ref.update_best_scores(pay_n, ra_n, new, pay_b, ra_b, b, 4, metric)
        out[f"scores.{metric}"] = sc.numpy()
        out[f"src.{metric}"] = ra[:, 0].numpy().astype(np.int64)
    return out


def pixel_cases(ref):
    """Pixel-space masked MSE: the reference has no pixel search, so its
    weighted_MSE (utils/similarity.py:174-192) is driven row by row with w = validity mask on NaN-zeroed
    tensors; weighted_MSE * D = sum m (q-x)^2 / sum m, the definition up to the 1e-5 in the normaliser."""
    out = {}
    seed = synth.BASE_SEED
    n, C, H, W, Q = 40, 5, 16, 16, 3
    x = synth.cutouts(n, C, H, W, seed=seed, stream=51)
    q = synth.cutouts(Q, C, H, W, seed=seed, stream=52, nan_frac=0.01, nan_chan_p=0.2)
    rng = np.random.Generator(np.random.PCG64([seed, 53]))
    qmask = (rng.random((Q, C, H, W)) < 0.6).astype(np.uint8)
    qmask[0] = 1                                    # query 0: no patch mask
    x[7] = np.nan                                   # a cutout with nothing valid
    out["meta"] = np.array([n, C, H, W, Q, seed, 51, 52, 53], dtype=np.int64)
    out["checksum"] = np.array(checksum(x, q, qmask))
    D = C * H * W
    ratio = np.zeros((Q, n), dtype=np.float64)      # sum m (q-x)^2 / sum m   from the reference
    msum = np.zeros((Q, n), dtype=np.float64)
    for qi in range(Q):
        qf = torch.from_numpy(q[qi].reshape(-1))
        for r in range(n):
            xf = torch.from_numpy(x[r].reshape(-1))
            m = (~torch.isnan(qf)) & (~torch.isnan(xf)) & torch.from_numpy(qmask[qi].reshape(-1) != 0)
            q0, x0 = torch.nan_to_num(qf, nan=0.0).double(), torch.nan_to_num(xf, nan=0.0).double()
            msum[qi, r] = float(m.sum())
            if m.any():
                v = ref.weighted_MSE(q0, x0[None, None], m.double())      # [1, 1]
                ratio[qi, r] = float(v[0, 0]) * D
    out["ratio"] = ratio
    out["msum"] = msum
    return out


def c1_cases(ref):
    """BASELINE config 1: similarity_search.py's call (similarity_search.py:169-171) with a random-init mim_1-shaped
    encoder (tests/stub_encoder.StubViT: timm is absent), 1k target cutouts, 10k bank cutouts, cosine, top-10, CPU."""
    from tests.stub_encoder import CutoutLoader, StubViT, c1_inputs
    out = {}
    bank, tgt, anchors = c1_inputs()
    model = StubViT(seed=0)
    with torch.no_grad():
        target_latent = torch.cat([model.forward_features(torch.from_numpy(tgt[s:s + 100]))[0] for s in range(0, len(tgt), 100)])
    out["checksum"] = np.array(checksum(bank[:64], tgt[:8]))
    out["anchors"] = anchors
    for mp, name in ((True, "maxpool"), (False, "patches")):
        res = quiet(ref.mae_simsearch, model, target_latent, CutoutLoader(bank, 64), "cpu", metric="cosine", combine="min",
                    use_weights=True, max_pool=mp, cls_token=False, nested_batches=False, n_save=10)
        out[f"scores.{name}"] = res[3].numpy()
        out[f"idx.{name}"] = res[2][:, 0].numpy().astype(np.int64)
    return out


def ingest_inputs(seed=synth.BASE_SEED):
    """Seeded inputs of the ingest fixtures (shared with tests/golden_inputs.py): cutouts with a central source of
    random brightness (so S/N spreads over the reference's default window [2, 7]) and a 3-band tile."""
    n, C, S = 48, 6, 64
    x = synth.cutouts(n, C, S, S, seed=seed, stream=41, nan_frac=0.0, nan_chan_p=0.1)     # missing bands only
    rng = np.random.Generator(np.random.PCG64([seed, 42]))
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32)
    blob = np.exp(-((yy - 31.5) ** 2 + (xx - 31.5) ** 2) / (2 * 3.0 ** 2)).astype(np.float32)
    amp = (rng.random((n, C)) * 12.0).astype(np.float32)
    x = (x + amp[:, :, None, None] * blob).astype(np.float32)
    x[5] = np.nan                      # a fully missing item
    x[3, 1, 30, 30] = np.nan           # one NaN pixel in the central region / in the surround: that channel's S/N is NaN
    x[4, 2, 2, 60] = np.nan
    x[9, 5] = 100.0                    # sixth channel is ignored by the nanmin over the first five
    tile = synth.cutouts(1, 3, 200, 173, seed=seed, stream=43, nan_frac=0.01, nan_chan_p=0.0)[0]
    tile[1, 100:, :] = np.nan
    tile = tile * 4.0                  # values below the -3 clip
    return x, tile


def ingest_cases(ref):
    """S/N pre-filter (utils/misc.py:119-163 + similarity_search.py:124-130), overlap coordinates and tile cutouts
    (utils/dataloaders.py:481-536) and extract_center (:685-700), by executing the reference's functions."""
    from oracle.ref_harness import load_reference_module
    misc = load_reference_module("utils/misc.py", "ref_misc")
    dl = load_reference_module("utils/dataloaders.py", "ref_dataloaders")
    out = {}
    x, tile = ingest_inputs()
    out["checksum"] = np.array(checksum(x, tile))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        snr = misc.calculate_snr(x, 8)
        mn = np.nanmin(snr[:, :5], axis=(1))                                     # similarity_search.py:127
    out["snr"] = snr.astype(np.float32)
    out["min_snr"] = mn.astype(np.float32)
    out["snr_range"] = np.array([2.0, 7.0])
    out["test_indices"] = np.where((mn > 2.0) & (mn < 7.0))[0].astype(np.int64)  # :130
    shapes = [(200, 173, 64, 0.4), (128, 128, 64, 0.5), (112, 112, 64, 0.25), (65, 64, 64, 0.0), (300, 90, 32, 0.6),
              (100, 100, 64, 0.4)]
    out["coord_cases"] = np.array(shapes, dtype=np.float64)
    for i, (H, W, size, ov) in enumerate(shapes):
        out[f"coords.{i}"] = np.array(dl.generate_overlap_coords((H, W), size, ov), dtype=np.int32).reshape(-1, 2)

    def pix_to_radec(h, w):            # an affine stand-in for the WCS call (argument order as the reference passes them)
        h, w = np.asarray(h, np.float64), np.asarray(w, np.float64)
        return 30.0 + 1e-3 * h + 1e-5 * w, -5.0 + 2e-3 * w - 1e-5 * h
    cut, ra_dec = dl.overlapping_cutouts(tile, 64, 0.4, pix_to_radec)
    cut = cut.copy()
    cut[cut < -3.0] = -3.0                                                       # utils/dataloaders.py:657-659
    out["tile_ra_dec"] = ra_dec.astype(np.float32)
    out["tile_cutouts_sha"] = np.array(hashlib.sha256(np.ascontiguousarray(cut.astype(np.float32)).tobytes()).hexdigest())
    out["tile_cutouts_first"] = cut[:2].astype(np.float32)
    big = synth.cutouts(3, 5, 96, 96, seed=synth.BASE_SEED, stream=44)
    out["center64_sha"] = np.array(hashlib.sha256(np.ascontiguousarray(
        np.stack([dl.extract_center(b, 64) for b in big]).astype(np.float32)).tobytes()).hexdigest())
    return out


def cli_cases(ref):
    """similarity_search.py's pipeline (:122-181) on small inputs, every numeric step by the reference's own code:
    calculate_snr + the S/N window (utils/misc.py, similarity_search.py:124-130), mae_latent (utils/eval_fns.py:72-140,
    no augmentation so that it is deterministic), mae_simsearch (:169-171).  The encoder is the mim_1-shaped stub; the
    loaders are in-memory equivalents of H5Dataset + DataLoader(shuffle=False) (clip at -3, zeros mask, [ra, dec])."""
    from oracle.ref_harness import load_reference_module
    from tests.golden_inputs import cli_inputs
    from tests.stub_encoder import StubViT
    misc = load_reference_module("utils/misc.py", "ref_misc")
    sys.path.insert(0, os.path.dirname(os.path.abspath(load_reference_module.__globals__["ref_path"]("utils/eval_fns.py"))))
    ev = load_reference_module("utils/eval_fns.py", "ref_eval_fns")
    inp = cli_inputs()
    out = {"checksum": np.array(checksum(inp["test"]["cutouts"], inp["target"]["cutouts"]))}
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        snr = np.nanmin(misc.calculate_snr(inp["test"]["cutouts"], 8)[:, :5], axis=(1))
    test_indices = np.where((snr > 2) & (snr < 7))[0]
    out["test_indices"] = test_indices.astype(np.int64)
    out["min_snr"] = snr.astype(np.float32)

    class Loader:
        def __init__(self, d, indices, bs):
            self.d, self.idx, self.bs = d, np.asarray(indices), bs

        def __len__(self):
            return (len(self.idx) + self.bs - 1) // self.bs

        def __iter__(self):
            for s in range(0, len(self.idx), self.bs):
                r = self.idx[s:s + self.bs]
                x = self.d["cutouts"][r].copy()
                x[x < -3.0] = -3.0
                t = torch.from_numpy(x)
                yield t, torch.zeros_like(t), torch.from_numpy(np.stack((self.d["ra"][r], self.d["dec"][r]), -1))

    model = torch.nn.DataParallel(StubViT(seed=0))          # the reference reads model.module.* (eval_fns.py:118)
    tgt_i = [1, 2, 4]
    target_latent, target_images = quiet(ev.mae_latent, model, Loader(inp["target"], tgt_i, 64), "cpu", return_images=True,
                                         apply_augmentations=False, num_augmentations=64, remove_cls=False)
    out["target_indices"] = np.array(tgt_i)
    out["target_features_sum"] = target_latent.double().sum(dim=(1, 2)).numpy()
    out["target_images_sha"] = np.array(hashlib.sha256(target_images.numpy().tobytes()).hexdigest())
    for mp, name in ((True, "maxpool"), (False, "patches")):
        imgs, lat, ra, sc = quiet(ref.mae_simsearch, model, target_latent, Loader(inp["test"], test_indices, 64), "cpu",
                                  metric="cosine", combine="min", use_weights=True, max_pool=mp, cls_token=False,
                                  nested_batches=False, n_save=12)
        out[f"test_scores.{name}"] = sc.numpy()
        out[f"test_ra_decs.{name}"] = ra.numpy()
        out[f"test_images_sum.{name}"] = imgs.double().nansum(dim=(1, 2, 3)).numpy()
        out[f"test_features_sum.{name}"] = lat.double().sum(dim=(1, 2)).numpy()
    return out


def c1_real_cases(ref):
    """BASELINE config 1 with the reference's OWN encoder: utils/mim_vit.py (unmodified, behind the test-only timm shim,
    tests/ref_encoder.py), random-init `mim_1` (seed 0), 1k target cutouts, 10k bank cutouts, cosine top-10 on the CPU
    through the reference's mae_simsearch exactly as similarity_search.py:169-171 calls it (use_weights=True)."""
    from tests.ref_encoder import build_mim1
    from tests.stub_encoder import CutoutLoader, c1_inputs
    out = {}
    bank, tgt, anchors = c1_inputs()
    model, _ = build_mim1("cpu", seed=0)
    with torch.no_grad():
        target_latent = torch.cat([model.module.forward_features(torch.from_numpy(tgt[s:s + 100]), reshape_out=False)[0]
                                   for s in range(0, len(tgt), 100)])
    out["checksum"] = np.array(checksum(bank[:64], tgt[:8]))
    out["anchors"] = anchors
    out["target_latent_sum"] = target_latent.double().sum(dim=(1, 2)).numpy()[:16]
    for mp, name in ((True, "maxpool"), (False, "patches")):
        res = quiet(ref.mae_simsearch, model, target_latent, CutoutLoader(bank, 64), "cpu", metric="cosine", combine="min",
                    use_weights=True, max_pool=mp, cls_token=False, nested_batches=False, n_save=10)
        out[f"scores.{name}"] = res[3].numpy()
        out[f"idx.{name}"] = res[2][:, 0].numpy().astype(np.int64)
    return out


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)   # deterministic reduction order for the fixtures
    only = sys.argv[1:]
    for name, fn in (("simsearch_small", simsearch_cases), ("simsearch_mim1_shape", mim1_shape_case),
                     ("short_bank", short_bank_case), ("compute_similarity", compute_similarity_cases),
                     ("update_best", update_best_cases), ("pixel_small", pixel_cases), ("c1_mim1_stub", c1_cases),
                     ("ingest", ingest_cases), ("cli_small", cli_cases),
                     ("c1_mim1_real", c1_real_cases)):
        if only and name not in only:
            continue
        if name in ("c1_mim1_stub", "c1_mim1_real"):
            torch.set_num_threads(os.cpu_count() or 1)      # the encoder pass over 11k cutouts
        data = fn(ref)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"wrote {path} ({os.path.getsize(path)} B, {len(data)} arrays)")


if __name__ == "__main__":
    main()

"""Rebuild the seeded inputs the golden fixtures were generated from (oracle/make_golden.py)."""
import hashlib
import os

import numpy as np

from sky_embeddings_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def checksum(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def simsearch_inputs(g):
    n, P, D, bs, k, seed, s_bank, s_tgt, copies = [int(v) for v in g["meta"]]
    bank = synth.latents(n, 1 + P, D, seed=seed, stream=s_bank)
    anchors = g["anchors"] if "anchors" in g else np.array([3])
    noise = float(g["noise"]) if "noise" in g else 0.3
    tgt = synth.target_group(bank, anchors, copies=copies, noise=noise, seed=seed, stream=s_tgt)
    assert checksum(bank, tgt) == str(g["checksum"]), "synthetic generator drifted from the fixtures"
    return bank, tgt, bs, k


def parse_simsearch_name(name):
    metric, mode, combine, uw = name.split(".")
    kw = dict(metric=metric, combine=combine, use_weights=(uw == "w"),
              max_pool=(mode == "maxpool"), cls_token=(mode == "cls"))
    return kw


def pixel_inputs(g):
    n, C, H, W, Q, seed, s_x, s_q, s_m = [int(v) for v in g["meta"]]
    x = synth.cutouts(n, C, H, W, seed=seed, stream=s_x)
    q = synth.cutouts(Q, C, H, W, seed=seed, stream=s_q, nan_frac=0.01, nan_chan_p=0.2)
    rng = np.random.Generator(np.random.PCG64([seed, s_m]))
    qmask = (rng.random((Q, C, H, W)) < 0.6).astype(np.uint8)
    qmask[0] = 1
    x[7] = np.nan
    assert checksum(x, q, qmask) == str(g["checksum"]), "synthetic generator drifted from the fixtures"
    return x, q, qmask


def ingest_inputs(g=None):
    """Cutouts with a central source + a 3-band tile, as oracle/make_golden.py::ingest_inputs built them."""
    seed = synth.BASE_SEED
    n, C, S = 48, 6, 64
    x = synth.cutouts(n, C, S, S, seed=seed, stream=41, nan_frac=0.0, nan_chan_p=0.1)
    rng = np.random.Generator(np.random.PCG64([seed, 42]))
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32)
    blob = np.exp(-((yy - 31.5) ** 2 + (xx - 31.5) ** 2) / (2 * 3.0 ** 2)).astype(np.float32)
    amp = (rng.random((n, C)) * 12.0).astype(np.float32)
    x = (x + amp[:, :, None, None] * blob).astype(np.float32)
    x[5] = np.nan
    x[3, 1, 30, 30] = np.nan
    x[4, 2, 2, 60] = np.nan
    x[9, 5] = 100.0
    tile = synth.cutouts(1, 3, 200, 173, seed=seed, stream=43, nan_frac=0.01, nan_chan_p=0.0)[0]
    tile[1, 100:, :] = np.nan
    tile = tile * 4.0
    if g is not None:
        assert checksum(x, tile) == str(g["checksum"]), "synthetic generator drifted from the fixtures"
    return x, tile


def tile_pix_to_radec(h, w):
    h, w = np.asarray(h, np.float64), np.asarray(w, np.float64)
    return 30.0 + 1e-3 * h + 1e-5 * w, -5.0 + 2e-3 * w - 1e-5 * h


def cli_inputs():
    """Small h5-shaped inputs of the CLI fixture: a 300-cutout test bank with central sources of varying brightness
    (so the S/N window keeps a subset) and a 6-cutout target file; ra / dec identify rows."""
    seed = synth.BASE_SEED
    n, C, S = 300, 5, 64
    x = synth.cutouts(n, C, S, S, seed=seed, stream=51, nan_frac=0.0, nan_chan_p=0.0)
    rng = np.random.Generator(np.random.PCG64([seed, 52]))
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32)
    blob = np.exp(-((yy - 31.5) ** 2 + (xx - 31.5) ** 2) / (2 * 3.0 ** 2)).astype(np.float32)
    amp = (2.0 + rng.random((n, C)) * 8.0).astype(np.float32)
    x = (x + amp[:, :, None, None] * blob).astype(np.float32)
    x *= 2.0                                             # some pixels below the -3 clip
    ra = (150.0 + 0.001 * np.arange(n)).astype(np.float32)
    dec = (2.0 - 0.002 * np.arange(n)).astype(np.float32)
    tgt_rows = np.array([11, 42, 77, 130, 201, 250])
    tgt = (x[tgt_rows] + 0.3 * rng.standard_normal((len(tgt_rows), C, S, S), dtype=np.float32)).astype(np.float32)
    tgt[np.isnan(tgt)] = 0.0
    return dict(test=dict(cutouts=x, ra=ra, dec=dec),
                target=dict(cutouts=tgt, ra=ra[tgt_rows].copy(), dec=dec[tgt_rows].copy()), tgt_rows=tgt_rows)

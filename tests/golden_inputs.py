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

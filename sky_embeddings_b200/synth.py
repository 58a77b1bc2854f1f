"""Seeded synthetic embeddings / cutouts shaped like the reference's data.

There is no network for the HSC datasets or ViT checkpoints, so tests and the bench use
synthetic tensors of the named shapes.  Embeddings mimic post-LayerNorm ViT features with
a per-feature spread so that the reference's inverse-variance weights
(/root/reference/utils/similarity.py:143-145) are non-trivial:

    x[r, d] = s_d * g[r, d] + m_d,   g ~ N(0,1),  s_d = exp(0.5 N(0,1)),  m_d ~ N(0, 0.5)

Two generators are provided:
  * numpy (PCG64, stable across versions) for small parity cases shared with the oracle;
  * torch-on-device, chunked and keyed on (seed, chunk index), for banks too large to come
    from the host (any row-sharding that is a multiple of CHUNK_ROWS yields the same bank).
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20240607
CHUNK_ROWS = 1 << 16


def feature_profile(D, seed=BASE_SEED):
    """Per-feature scale s_d and offset m_d (shared by bank and queries)."""
    rng = np.random.Generator(np.random.PCG64([seed, 0xFEA7]))
    s = np.exp(0.5 * rng.standard_normal(D)).astype(np.float32)
    m = (0.5 * rng.standard_normal(D)).astype(np.float32)
    return s, m


def latents(n, tokens, D, seed=BASE_SEED, stream=1, dtype=np.float32):
    """[n, tokens, D] synthetic encoder output (token 0 plays the cls token)."""
    s, m = feature_profile(D, seed)
    rng = np.random.Generator(np.random.PCG64([seed, stream]))
    g = rng.standard_normal((n, tokens, D), dtype=np.float32)
    return (g * s + m).astype(dtype)


def target_group(bank_latents, anchor_rows, copies, noise=0.1, seed=BASE_SEED, stream=2):
    """A target group like the reference's augmented targets
    (/root/reference/similarity_search.py:160-162: each target + 64 augmented copies):
    ``copies`` noisy versions of each anchor bank item.  [len(anchor_rows)*copies, tokens, D]."""
    rng = np.random.Generator(np.random.PCG64([seed, stream]))
    base = bank_latents[np.asarray(anchor_rows)]
    rep = np.repeat(base, copies, axis=0)
    return (rep + noise * rng.standard_normal(rep.shape, dtype=np.float32)).astype(np.float32)


def cutouts(n, C=5, H=64, W=64, seed=BASE_SEED, stream=3, nan_frac=0.02, nan_chan_p=0.05):
    """[n, C, H, W] f32 pixel cutouts ~ N(0,1) clipped at -3
    (/root/reference/utils/dataloaders.py:294-295) with NaN pixels and whole-channel NaNs
    (missing bands, /root/reference/utils/dataloaders.py:442-445)."""
    rng = np.random.Generator(np.random.PCG64([seed, stream]))
    x = rng.standard_normal((n, C, H, W), dtype=np.float32)
    np.maximum(x, -3.0, out=x)
    if nan_frac > 0:
        x[rng.random((n, C, H, W), dtype=np.float32) < nan_frac] = np.nan
    if nan_chan_p > 0:
        x[rng.random((n, C)) < nan_chan_p] = np.nan
    return x


# ------------------------------------------------------------------------------------------
# device-side generation for full-size banks (bench / large GPU tests)
# ------------------------------------------------------------------------------------------

def device_bank_chunk(chunk_index, rows, D, device, seed=BASE_SEED, dtype=None):
    """Rows [chunk_index*CHUNK_ROWS, +rows) of the full-size synthetic bank, generated on
    ``device`` with a torch Philox generator keyed on (seed, chunk_index)."""
    import torch
    s, m = feature_profile(D, seed)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 1000003 + int(chunk_index))
    g = torch.randn((rows, D), generator=gen, device=device, dtype=torch.float32)
    x = g * torch.from_numpy(s).to(device) + torch.from_numpy(m).to(device)
    return x if dtype is None else x.to(dtype)


def device_bank_rows(row0, rows, D, device, seed=BASE_SEED):
    """Arbitrary row range of the full-size bank (row0 must be CHUNK_ROWS aligned)."""
    import torch
    assert row0 % CHUNK_ROWS == 0, "row0 must be a multiple of CHUNK_ROWS"
    out = []
    done = 0
    while done < rows:
        n = min(CHUNK_ROWS, rows - done)
        out.append(device_bank_chunk((row0 + done) // CHUNK_ROWS, n, D, device, seed))
        done += n
    return torch.cat(out) if len(out) > 1 else out[0]


def planted_queries(bank_rows_fn, n_total, Q, D, device, noise=0.1, seed=BASE_SEED):
    """Q queries = bank rows at a fixed stride + N(0, noise^2): the planted row is the known
    top-1 answer.  Returns (queries[Q, D] f32, planted_rows[Q])."""
    import torch
    stride = max(n_total // Q, 1)
    rows = [(q * stride + stride // 2) % n_total for q in range(Q)]
    qs = []
    for r in rows:
        c, off = divmod(r, CHUNK_ROWS)
        chunk_rows = min(CHUNK_ROWS, n_total - c * CHUNK_ROWS)
        qs.append(bank_rows_fn(c, chunk_rows)[off])
    q = torch.stack(qs).to(torch.float32)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 7919 + 17)
    q = q + noise * torch.randn(q.shape, generator=gen, device=device, dtype=torch.float32)
    return q, rows

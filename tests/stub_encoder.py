"""A ViT-shaped, random-init stand-in for the reference's ``mim_1`` encoder (configs/mim_1.ini:21-28: 64 px
images, 8 px patches, 5 channels, embed dim 768; utils/mim_vit.py:381-438 ``forward_features``).

timm (and astropy / h5py) are not installed here, so the reference's own ViT cannot be instantiated
(SURVEY.md section 8(c)); the search path only consumes the latents an encoder emits, so BASELINE config 1
is exercised with this deterministic stub on both sides: the reference's mae_simsearch on the CPU
(oracle/make_golden.py -> tests/golden/c1_mim1_stub.npz) and the CUDA path on the GPU.
Test infrastructure only.
"""
import math

import torch


class StubViT(torch.nn.Module):
    num_extra_tokens = 1
    attn_pool = False

    def __init__(self, seed=0, channels=5, img=64, patch=8, dim=768):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.patch, self.n_patch, self.dim = patch, (img // patch) ** 2, dim
        pdim = channels * patch * patch
        self.proj = torch.nn.Parameter(torch.randn(pdim, dim, generator=g) / math.sqrt(pdim), requires_grad=False)
        self.mix = torch.nn.Parameter(torch.randn(dim, dim, generator=g) / math.sqrt(dim), requires_grad=False)
        self.cls = torch.nn.Parameter(0.02 * torch.randn(1, 1, dim, generator=g), requires_grad=False)
        self.pos = torch.nn.Parameter(0.02 * torch.randn(1, 1 + self.n_patch, dim, generator=g), requires_grad=False)

    def forward_features(self, samples, ra_dec=None, mask=None, reshape_out=False):
        B, C, H, W = samples.shape
        # missing bands / pixels are NaN; the reference's encoder replaces them with its (zero-initialised) mask values
        # before patch embedding (utils/mim_vit.py:388-392)
        samples = torch.where(torch.isnan(samples), torch.zeros_like(samples), samples)
        p = self.patch
        x = samples.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, self.n_patch, C * p * p)
        x = x @ self.proj
        x = torch.cat([self.cls.expand(B, -1, -1), x], dim=1) + self.pos
        x = x + torch.tanh(x @ self.mix)
        x = torch.nn.functional.layer_norm(x, (self.dim,))
        return x, None, None


class CutoutLoader:
    """Flat loader over a cutout bank: (samples [B,C,H,W], mask, ra_dec [B,2]); ra_dec[:,0] = bank row."""

    def __init__(self, cutouts, batch_size):
        self.x, self.bs = torch.as_tensor(cutouts), batch_size

    def __len__(self):
        return (self.x.shape[0] + self.bs - 1) // self.bs

    def __iter__(self):
        n = self.x.shape[0]
        for s in range(0, n, self.bs):
            e = min(n, s + self.bs)
            ra = torch.zeros((e - s, 2))
            ra[:, 0] = torch.arange(s, e, dtype=torch.float32)
            yield self.x[s:e], torch.zeros(e - s), ra


def c1_inputs(n_bank=10000, n_target=1000, seed_stream=(81, 82)):
    """BASELINE config 1 data: 10k bank cutouts, a target group of 1k noisy copies of two bank cutouts."""
    import numpy as np
    from sky_embeddings_b200 import synth
    bank = synth.cutouts(n_bank, 5, 64, 64, stream=seed_stream[0], nan_frac=0.0, nan_chan_p=0.0)
    rng = np.random.Generator(np.random.PCG64([synth.BASE_SEED, seed_stream[1]]))
    anchors = np.array([123, 4567])
    tgt = np.repeat(bank[anchors], n_target // 2, axis=0) + 0.5 * rng.standard_normal((n_target, 5, 64, 64), dtype=np.float32)
    return bank, tgt.astype(np.float32), anchors

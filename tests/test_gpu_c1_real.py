"""BASELINE config 1 with the reference's own `mim_1` encoder (utils/mim_vit.py, unmodified, random init, behind the
test-only timm shim): 1k target cutouts, 10k bank cutouts, cosine top-10, use_weights=True -- the call of
similarity_search.py:169-171.  The golden (tests/golden/c1_mim1_real.npz) is the reference's mae_simsearch with that
model on the CPU; here the same model runs on the GPU in front of the drop-in mae_simsearch and of the resident route.
The encoder runs on different hardware on the two sides (fp32 ViT-Base, 12 blocks), so the comparison with the golden
uses 1e-4 absolute on cosine scores; against the CPU oracle fed the SAME GPU latents it is the fp32 1e-5."""
import numpy as np
import pytest
import torch

from oracle import sky_oracle as O
from tests import golden_inputs as G

pytestmark = pytest.mark.gpu


def test_c1_reference_mim1_encoder_end_to_end():
    from tests import ref_encoder
    if not ref_encoder.available():
        pytest.fail("oracle/_ref is missing: run __graft_entry__.build() in the build container before shipping")
    from sky_embeddings_b200 import bank_from_loader, resident_simsearch
    from sky_embeddings_b200 import similarity as S
    from tests.stub_encoder import CutoutLoader, c1_inputs
    dev = torch.device("cuda:0")
    g = G.load("c1_mim1_real")
    bank, tgt, anchors = c1_inputs()
    assert G.checksum(bank[:64], tgt[:8]) == str(g["checksum"])
    model, _cfg = ref_encoder.build_mim1("cpu", seed=0)
    model = model.to(dev)
    enc = model.module
    with torch.no_grad():
        target_latent = torch.cat([enc.forward_features(torch.from_numpy(tgt[s:s + 250]).to(dev), reshape_out=False)[0]
                                   for s in range(0, len(tgt), 250)])
        bank_latent = torch.cat([enc.forward_features(torch.from_numpy(bank[s:s + 500]).to(dev), reshape_out=False)[0]
                                 for s in range(0, len(bank), 500)])
    assert np.allclose(target_latent.double().sum(dim=(1, 2)).cpu().numpy()[:16], g["target_latent_sum"], rtol=1e-3, atol=0.5)
    for mp, name in ((True, "maxpool"), (False, "patches")):
        smp, lat, ra, sc = S.mae_simsearch(model, target_latent, CutoutLoader(bank, 64), dev, metric="cosine", combine="min",
                                           use_weights=True, max_pool=mp, cls_token=False, nested_batches=False, n_save=10)
        idx = ra[:, 0].cpu().numpy().astype(np.int64)
        # (1) against the reference's own run (encoder on the CPU there)
        ok, msg = O.check_topk_parity(sc.cpu().numpy(), idx, g[f"scores.{name}"], g[f"idx.{name}"], 1e-4, scale=1.0)
        assert ok, f"mae_simsearch vs reference golden, {name}: {msg}"
        assert set(idx[:2].tolist()) == set(anchors.tolist()), "the two planted anchors lead"
        # (2) against the CPU oracle on the latents this GPU produced: the search itself at fp32 tolerance
        ref_s, ref_i, *_ = O.simsearch(target_latent.cpu().numpy(), bank_latent.cpu().numpy(), 64, 10, metric="cosine",
                                       combine="min", use_weights=True, max_pool=mp)
        ok, msg = O.check_topk_parity(sc.cpu().numpy(), idx, ref_s, ref_i, 1e-5, scale=1.0)
        assert ok, f"mae_simsearch vs oracle on the same latents, {name}: {msg}"
        assert np.array_equal(smp.cpu().numpy(), bank[idx]) and lat.shape == (10, 65, 768)
        # (3) resident route: one encoding pass into a device-resident bank, one search
        rb, rra, _ = bank_from_loader(model, CutoutLoader(bank, 64), dev, max_pool=mp, bank_dtype="fp32", n_items=len(bank))
        _, ridx, _, rsc = resident_simsearch(rb, target_latent, rra, None, 1, 10, "cosine", "min", True, mp, False)
        ok, msg = O.check_topk_parity(rsc.cpu().numpy(), ridx.cpu().numpy(), ref_s, ref_i, 1e-5, scale=1.0)
        assert ok, f"resident route, {name}: {msg}"
        rb.close()

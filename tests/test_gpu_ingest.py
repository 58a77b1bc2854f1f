"""GPU parity of the ingest kernels (csrc/pixel_prep.cu) against fixtures produced by executing the reference's own
functions (oracle/make_golden.py::ingest_cases): calculate_snr + the S/N window of similarity_search.py:124-130,
overlapping_cutouts with the edge rules of generate_overlap_coords, extract_center; and the h5 / tile loaders end to
end through the drop-in mae_simsearch."""
import hashlib

import numpy as np
import pytest
import torch

from tests import golden_inputs as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _close(a, b, rel=1e-5, abs_=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    ok = both_nan | (np.abs(a - b) <= abs_ + rel * np.abs(b))
    return bool(ok.all()), int((~ok).sum())


def test_snr_matches_reference_golden(dev):
    from sky_embeddings_b200 import ingest
    g = G.load("ingest")
    x, _ = G.ingest_inputs(g)
    snr, mn = ingest.snr_device(torch.from_numpy(x).to(dev), n_central_pix=8, n_min_channels=5)
    ok, bad = _close(snr.cpu().numpy(), g["snr"])
    assert ok, f"{bad} S/N values differ from calculate_snr"
    ok, bad = _close(mn.cpu().numpy(), g["min_snr"])
    assert ok, f"{bad} nanmin values differ"
    assert np.isnan(mn[5].item()), "an all-NaN item has no S/N"
    # a different central size, against the restated definition (fp64)
    snr4, _ = ingest.snr_device(torch.from_numpy(x).to(dev), n_central_pix=20, n_min_channels=5)
    xs = x.astype(np.float64)
    m = np.ones((64, 64), bool); m[22:42, 22:42] = False
    want = xs[:, :, 22:42, 22:42].mean((2, 3)) / (xs[:, :, m].std(2) + 1e-8)
    ok, bad = _close(snr4.cpu().numpy(), want, rel=2e-5, abs_=2e-6)
    assert ok, bad


def test_snr_window_selects_the_reference_rows(dev, tmp_path):
    from sky_embeddings_b200 import h5lite, ingest
    g = G.load("ingest")
    x, _ = G.ingest_inputs(g)
    n = len(x)
    p = h5lite.write_h5(str(tmp_path / "bank.h5"), dict(cutouts=x, ra=np.arange(n, dtype="f"), dec=np.zeros(n, "f")))
    src = ingest.H5Cutouts(p)
    idx, mn = src.select_snr(dev, g["snr_range"], n_central_pix=8, batch_size=17)     # ragged slices of the file
    want = g["test_indices"]
    # rows whose S/N sits within tolerance of a window edge may fall on either side
    edge = np.zeros(n, bool)
    for b in g["snr_range"]:
        edge |= np.abs(g["min_snr"] - b) <= 1e-4 * abs(b)
    assert set(idx.tolist()) - set(np.where(edge)[0].tolist()) == set(want.tolist()) - set(np.where(edge)[0].tolist())
    assert np.all(np.diff(idx) > 0), "bank row order = ascending test_indices (similarity_search.py:130, shuffle=False)"
    # pixel bank of the selected rows: row i is file row idx[i], clipped
    pb = src.pixel_bank(dev, indices=idx)
    q = np.maximum(x[idx[3]], -3.0)[None]
    sc, ix = pb.search(torch.from_numpy(q).to(dev), None, k=1)
    assert int(ix[0, 0]) == 3 and float(sc[0, 0]) < 1e-6
    pb.close()
    src.close()


def test_tile_cutouts_match_reference_golden(dev):
    from sky_embeddings_b200 import ingest
    g = G.load("ingest")
    _, tile = G.ingest_inputs(g)
    coords = ingest.generate_overlap_coords(tile.shape[1:], 64, 0.4)
    cut = ingest.tile_cutouts_device(torch.from_numpy(tile).to(dev), coords, 64, pixel_min=-3.0)
    got = cut.cpu().numpy()
    assert np.array_equal(got[:2], g["tile_cutouts_first"], equal_nan=True)
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(g["tile_cutouts_sha"]), \
        "gathered cutouts are not bit-identical to overlapping_cutouts + clip"
    with pytest.raises(ValueError):
        ingest.tile_cutouts_device(torch.from_numpy(tile).to(dev), np.array([[150, 0]], np.int32), 64)


def test_center_clip_matches_reference_golden(dev):
    from sky_embeddings_b200 import ingest, synth
    g = G.load("ingest")
    big = synth.cutouts(3, 5, 96, 96, seed=synth.BASE_SEED, stream=44)
    got = ingest.center_clip_device(torch.from_numpy(big).to(dev), 64, pixel_min=None).cpu().numpy()
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(g["center64_sha"])
    clipped = ingest.center_clip_device(torch.from_numpy(big).to(dev), 64, pixel_min=-1.0, pixel_max=0.5).cpu().numpy()
    want = big[:, :, 16:80, 16:80].copy()
    want[want < -1.0] = -1.0
    want[want > 0.5] = 0.5
    assert np.array_equal(clipped, want, equal_nan=True)


def test_h5_loader_device_mode_equals_host_mode(dev, tmp_path):
    from sky_embeddings_b200 import h5lite, ingest
    rng = np.random.default_rng(8)
    n = 37
    x = (rng.standard_normal((n, 5, 80, 80)) * 4).astype(np.float32)
    x[6, 3] = np.nan
    p = h5lite.write_h5(str(tmp_path / "b.h5"), dict(cutouts=x, ra=rng.random(n).astype("f"), dec=rng.random(n).astype("f")))
    src = ingest.H5Cutouts(p, img_size=64)
    idx = np.array([0, 3, 4, 9, 10, 11, 30, 36, 6])
    host = list(src.loader(idx, batch_size=4))
    devb = list(src.loader(idx, batch_size=4, device=dev))
    assert len(host) == len(devb) == 3
    for (hx, hm, hr), (dx, dm, dr) in zip(host, devb):
        assert dx.is_cuda and np.array_equal(dx.cpu().numpy(), hx.numpy(), equal_nan=True)
        assert np.array_equal(dr.cpu().numpy(), hr.numpy()) and dm.shape == dx.shape
    src.close()


def test_tile_stream_through_mae_simsearch(dev):
    """The sky_sim_search.py route: nested per-tile batches from the tile loader (device mode, double-buffered
    uploads) through the drop-in mae_simsearch give what the host-mode loader gives through the same call, and the
    best match of a query cut from a tile is that cutout."""
    from sky_embeddings_b200 import ingest, synth
    from sky_embeddings_b200.similarity import mae_simsearch
    from tests.stub_encoder import StubViT
    tiles = []
    for t in range(3):
        tile = synth.cutouts(1, 5, 256, 256, seed=synth.BASE_SEED, stream=60 + t, nan_frac=0.0, nan_chan_p=0.0)[0]

        def p2r(h, w, t=t):
            return np.asarray(h, np.float64) + 1000.0 * t, np.asarray(w, np.float64)
        tiles.append((tile, p2r))
    model = StubViT(seed=0).to(dev)
    coords = ingest.generate_overlap_coords((256, 256), 64, 0.4)
    h0, w0 = coords[17]
    tgt_img = torch.from_numpy(tiles[1][0][None, :, h0:h0 + 64, w0:w0 + 64].copy()).to(dev)
    with torch.no_grad():
        target_latent = model.forward_features(tgt_img)[0]
    out = {}
    for mode, device in (("host", None), ("device", dev)):
        ld = ingest.TileLoader(tiles, batch_size=8, img_size=64, overlap=0.4, device=device)
        res = mae_simsearch(model, target_latent, ld, dev, metric="cosine", combine="min", use_weights=False,
                            max_pool=True, cls_token=False, nested_batches=True, n_save=5)
        out[mode] = [r.cpu().numpy() for r in res]
    for a, b in zip(out["host"], out["device"]):
        assert np.allclose(a, b, rtol=1e-5, atol=1e-6, equal_nan=True)
    ra_dec = out["device"][2]
    assert ra_dec[0, 0] == h0 + 32 + 1000.0 and ra_dec[0, 1] == w0 + 32, "top hit is the cutout the target was cut from"

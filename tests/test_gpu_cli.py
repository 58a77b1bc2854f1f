"""The command-line driver against the reference's pipeline (golden: oracle/make_golden.py::cli_cases executes the
reference's calculate_snr, mae_latent and mae_simsearch on the same small h5-shaped inputs)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from tests import golden_inputs as G

pytestmark = pytest.mark.gpu

KEYS = ["test_ra_decs", "test_scores", "target_images", "target_features", "test_images", "test_features"]


def stub_factory(config, device):
    """--encoder hook: the mim_1-shaped stub, wrapped like the reference wraps its model (nn.DataParallel, .module)."""
    from tests.stub_encoder import StubViT

    class Wrapped(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.module = StubViT(seed=0)
    return Wrapped().to(device)


def _write_inputs(tmp_path):
    from sky_embeddings_b200 import h5lite
    inp = G.cli_inputs()
    data = tmp_path / "data"
    data.mkdir()
    h5lite.write_h5(str(data / "tests_bank.h5"), inp["test"])
    h5lite.write_h5(str(data / "targets.h5"), inp["target"])
    return inp, data


@pytest.mark.parametrize("mp,name", [("True", "maxpool"), ("False", "patches")])
@pytest.mark.parametrize("resident", [False, True])
def test_cli_matches_reference_pipeline(tmp_path, mp, name, resident):
    from sky_embeddings_b200 import search
    g = G.load("cli_small")
    inp, data = _write_inputs(tmp_path)
    assert G.checksum(inp["test"]["cutouts"], inp["target"]["cutouts"]) == str(g["checksum"])
    argv = ["stubmodel", "-tgt_fn", "targets.h5", "-tst_fn", "tests_bank.h5", "-tgt_i", "[1,2,4]", "-aug", "False",
            "-mp", mp, "-ct", "False", "-snr", "[2,7]", "-bs", "64", "-m", "cosine", "-c", "min", "-ns", "12",
            "-dd", str(data), "--results-dir", str(tmp_path / "results"), "--encoder", "tests.test_gpu_cli:stub_factory"]
    if resident:
        argv.append("--resident")
    out = search.main(argv)
    # the reference's file name and keys (similarity_search.py:178-181)
    assert os.path.basename(out) == "stubmodel_targets_simsearch_results_f.npz"
    r = np.load(out)
    assert sorted(r.files) == sorted(KEYS)
    assert r["test_ra_decs"].shape == (12, 2) and r["test_scores"].shape == (12,)
    assert r["target_images"].shape == (3, 5, 64, 64) and r["target_features"].shape == (3, 65, 768)
    assert r["test_images"].shape == (12, 5, 64, 64) and r["test_features"].shape == (12, 65, 768)
    # targets: the reference's mae_latent
    assert hashlib.sha256(r["target_images"].tobytes()).hexdigest() == str(g["target_images_sha"])
    assert np.allclose(r["target_features"].astype(np.float64).sum((1, 2)), g["target_features_sum"], rtol=1e-4, atol=1e-2)
    # search results: scores within fp32 tolerance, winners identical (identified by their ra / dec)
    assert np.allclose(r["test_scores"], g[f"test_scores.{name}"], rtol=2e-5, atol=2e-6)
    gaps = np.abs(np.diff(g[f"test_scores.{name}"]))
    if gaps.min() > 1e-5:
        assert np.array_equal(r["test_ra_decs"], g[f"test_ra_decs.{name}"])
    else:
        assert {tuple(x) for x in r["test_ra_decs"].tolist()} == {tuple(x) for x in g[f"test_ra_decs.{name}"].tolist()}
    assert np.allclose(np.nansum(r["test_images"].astype(np.float64), axis=(1, 2, 3)), g[f"test_images_sum.{name}"], rtol=1e-6, atol=1e-3)
    assert np.allclose(r["test_features"].astype(np.float64).sum((1, 2)), g[f"test_features_sum.{name}"], rtol=1e-4, atol=1e-2)


def test_cli_tile_mode_writes_the_sky_sim_search_layout(tmp_path):
    """sky_sim_search.py route: tiles -> overlapping cutouts -> nested batches; result file without the `_f`."""
    from sky_embeddings_b200 import search, synth
    inp, data = _write_inputs(tmp_path)
    tiles = tmp_path / "tiles"
    tiles.mkdir()
    for t in range(2):
        np.save(tiles / f"tile{t}.npy", synth.cutouts(1, 5, 200, 200, stream=70 + t, nan_frac=0.0, nan_chan_p=0.0)[0])
    out = search.main(["stubmodel", "-tgt_fn", "targets.h5", "-tst_dirs", str(tiles), "-tgt_i", "[0,3]", "-aug", "True",
                       "-mp", "True", "-bs", "8", "-ns", "20", "-dd", str(data), "--results-dir", str(tmp_path / "results"),
                       "--encoder", "tests.test_gpu_cli:stub_factory"])
    assert os.path.basename(out) == "stubmodel_targets_simsearch_results.npz"
    r = np.load(out)
    assert sorted(r.files) == sorted(KEYS)
    assert r["target_images"].shape == (2 * 65, 5, 64, 64), "each target is followed by its 64 augmented copies"
    assert r["test_scores"].shape == (20,) and np.all(np.diff(r["test_scores"]) <= 0) and np.isfinite(r["test_scores"]).all()


def test_cli_tile_mode_reads_fits_tiles(tmp_path):
    """-tst_dirs over a directory of per-band FITS files (the reference's naming): patches are grouped, a missing band
    becomes a NaN plane, ra / dec come from the header's TAN WCS, and the best match of a target cut from a tile is found
    at the sky position of that cutout."""
    from sky_embeddings_b200 import fitslite, h5lite, ingest, search, synth
    hdr = {"CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CRVAL1": 150.25, "CRVAL2": 2.5, "CRPIX1": 101.0, "CRPIX2": 81.0,
           "CD1_1": -4.6e-5, "CD1_2": 0.0, "CD2_1": 0.0, "CD2_2": 4.6e-5}
    tiles = tmp_path / "tiles"
    tiles.mkdir()
    planes = synth.cutouts(1, 5, 200, 200, stream=90, nan_frac=0.0, nan_chan_p=0.0)[0]
    for i, b in enumerate("GRIZY"):
        fitslite.write_image(str(tiles / f"calexp-HSC-{b}-9813-4,4.fits"), planes[i], hdr)
    coords = ingest.generate_overlap_coords((200, 200), 64, 0.4)
    h0, w0 = coords[5]
    # two slightly noisy copies of the cutout: a single un-augmented target has no spread, and the reference's weights
    # 1 / std^2 (utils/similarity.py:143) are NaN then
    rng = np.random.default_rng(0)
    cut = planes[:, h0:h0 + 64, w0:w0 + 64]
    tgt = (np.stack([cut, cut]) + 0.01 * rng.standard_normal((2, 5, 64, 64))).astype(np.float32)
    data = tmp_path / "data"
    data.mkdir()
    h5lite.write_h5(str(data / "targets.h5"), dict(cutouts=tgt, ra=np.zeros(2, "f"), dec=np.zeros(2, "f")))
    out = search.main(["stubmodel", "-tgt_fn", "targets.h5", "-tst_dirs", str(tiles), "-tgt_i", "[0,1]", "-aug", "False", "-mp", "True",
                       "-bs", "4", "-ns", "5", "-dd", str(data), "--results-dir", str(tmp_path / "results"),
                       "--encoder", "tests.test_gpu_cli:stub_factory"])
    r = np.load(out)
    ra, dec = fitslite.TanWcs(hdr).all_pix2world([h0 + 32], [w0 + 32], 0)        # argument order as the reference calls it
    assert np.allclose(r["test_ra_decs"][0], [ra[0], dec[0]], atol=1e-4), "top hit is the cutout the target was cut from"
    assert np.array_equal(r["test_images"][0], np.maximum(cut, -3.0))

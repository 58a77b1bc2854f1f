"""CPU tests of the FITS tile reader and its TAN(-SIP) pixel -> sky transform (sky_embeddings_b200/fitslite.py).
astropy is not in this image, so the transform is checked against its defining identities and an independent inverse
(parity unpinned against astropy, as the module says); the file-name grouping and missing-band handling follow the
reference's find_HSC_bands / load_fits_bands (utils/dataloaders.py:330-437)."""
import numpy as np
import pytest

from sky_embeddings_b200 import fitslite as F

HDR = {"CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CRVAL1": 150.25, "CRVAL2": 2.5, "CRPIX1": 101.0, "CRPIX2": 81.0,
       "CD1_1": -4.6e-5, "CD1_2": 1.0e-7, "CD2_1": 2.0e-7, "CD2_2": 4.6e-5}


def test_image_round_trip_and_header(tmp_path):
    rng = np.random.default_rng(1)
    img = rng.standard_normal((160, 200)).astype(np.float32)
    img[3, 5] = np.nan
    p = F.write_image(str(tmp_path / "calexp-HSC-G-9813-4,4.fits"), img, dict(HDR, OBJECT="a 'quoted' name"))
    hdus = F.read_hdus(p)
    assert len(hdus) == 2 and hdus[0][0]["NAXIS"] == 0 and hdus[1][0]["XTENSION"].strip() == "IMAGE"
    data, hdr = F.read_image(p, 1)
    assert data.shape == (160, 200) and data.dtype == np.dtype(">f4")
    assert np.array_equal(np.asarray(data, np.float32), img, equal_nan=True)
    assert hdr["CRVAL1"] == 150.25 and hdr["CTYPE1"] == "RA---TAN" and hdr["NAXIS1"] == 200
    with pytest.raises(F.FitsError):
        F.read_image(p, 2)
    raw = open(p, "rb").read()
    (tmp_path / "cut.fits").write_bytes(raw[: len(raw) - 2 * F.BLOCK])
    with pytest.raises(F.FitsError):
        F.read_image(str(tmp_path / "cut.fits"), 1)
    (tmp_path / "junk.fits").write_bytes(b"x" * 5000)
    with pytest.raises(F.FitsError):
        F.read_hdus(str(tmp_path / "junk.fits"))


def _inverse_tan(w, ra, dec):
    """Independent inverse (sky -> pixel, 0-based) by the standard gnomonic formulas, no SIP."""
    a, d = np.deg2rad(ra), np.deg2rad(dec)
    a0, d0 = np.deg2rad(w.crval[0]), np.deg2rad(w.crval[1])
    cosc = np.sin(d0) * np.sin(d) + np.cos(d0) * np.cos(d) * np.cos(a - a0)
    xi = np.cos(d) * np.sin(a - a0) / cosc
    eta = (np.cos(d0) * np.sin(d) - np.sin(d0) * np.cos(d) * np.cos(a - a0)) / cosc
    uv = np.linalg.solve(w.cd, np.rad2deg(np.stack([xi, eta])))
    return uv[0] + w.crpix[0] - 1, uv[1] + w.crpix[1] - 1


def test_tan_wcs_identities():
    w = F.TanWcs(HDR)
    ra, dec = w.all_pix2world(HDR["CRPIX1"] - 1, HDR["CRPIX2"] - 1, 0)          # the reference pixel, 0-based
    assert abs(ra - 150.25) < 1e-12 and abs(dec - 2.5) < 1e-12
    rng = np.random.default_rng(2)
    x, y = rng.uniform(0, 4000, 50), rng.uniform(0, 4000, 50)
    ra, dec = w.all_pix2world(x, y, 0)
    # angular distance from the reference point = arctan of the tangent-plane radius
    u, v = x + 1 - w.crpix[0], y + 1 - w.crpix[1]
    r = np.hypot(*(w.cd @ np.stack([u, v])))
    a0, d0 = np.deg2rad(150.25), np.deg2rad(2.5)
    cosc = np.sin(d0) * np.sin(np.deg2rad(dec)) + np.cos(d0) * np.cos(np.deg2rad(dec)) * np.cos(np.deg2rad(ra) - a0)
    assert np.allclose(np.rad2deg(np.arccos(np.clip(cosc, -1, 1))), np.rad2deg(np.arctan(np.deg2rad(r))), rtol=0, atol=2e-9)
    xi, yi = _inverse_tan(w, ra, dec)
    assert np.allclose(xi, x, atol=1e-6) and np.allclose(yi, y, atol=1e-6)
    # north is +eta, east is +RA: one pixel up raises dec by CD2_2, one pixel right lowers RA by |CD1_1| / cos(dec)
    ra1, dec1 = w.all_pix2world(100.0, 81.0, 0)
    assert abs((dec1 - 2.5) - 4.6e-5) < 1e-9
    ra2, _ = w.all_pix2world(101.0, 80.0, 0)
    assert abs((ra2 - 150.25) - (-4.6e-5 / np.cos(np.deg2rad(2.5)))) < 1e-8
    # origin 1 shifts by one pixel; PC + CDELT form gives the same answers; SIP adds its polynomial before the CD matrix
    assert np.allclose(w.all_pix2world(x + 1, y + 1, 1), (ra, dec))
    h2 = {k: v for k, v in HDR.items() if not k.startswith("CD")}
    h2.update(CDELT1=-4.6e-5, CDELT2=4.6e-5, PC1_1=1.0, PC1_2=1.0e-7 / -4.6e-5, PC2_1=2.0e-7 / 4.6e-5, PC2_2=1.0)
    assert np.allclose(F.TanWcs(h2).all_pix2world(x, y, 0), (ra, dec), rtol=0, atol=1e-10)
    hs = dict(HDR, CTYPE1="RA---TAN-SIP", CTYPE2="DEC--TAN-SIP", A_ORDER=2, B_ORDER=2, A_2_0=1e-7, B_1_1=-2e-7)
    ras, decs = F.TanWcs(hs).all_pix2world(x, y, 0)
    ras2, decs2 = w.all_pix2world(x + 1e-7 * u ** 2, y - 2e-7 * u * v, 0)
    assert np.allclose(ras, ras2, atol=1e-11) and np.allclose(decs, decs2, atol=1e-11)
    with pytest.raises(F.FitsError):
        F.TanWcs(dict(HDR, CTYPE1="GLON-CAR"))


def test_band_grouping_and_missing_bands(tmp_path):
    rng = np.random.default_rng(3)
    tiles = {}
    for band in "GRIZ":                     # Y is missing for patch 4,4
        img = rng.standard_normal((96, 128)).astype(np.float32)
        tiles[band] = img
        F.write_image(str(tmp_path / f"calexp-HSC-{band}-9813-4,4.fits"), img, HDR)
    F.write_image(str(tmp_path / "calexp-HSC-G-9813-4,5.fits"), tiles["G"], HDR)     # a patch with one band only
    F.write_image(str(tmp_path / "HSC-G-9813-4,4.fits"), tiles["G"], HDR)            # not a calexp file
    bands = ["G", "R", "I", "Z", "Y"]
    groups = F.find_tile_bands([str(tmp_path)], bands, min_bands=4, use_calexp=True)
    assert len(groups) == 1 and groups[0][4] == "None" and all(g.endswith(f"calexp-HSC-{b}-9813-4,4.fits") for g, b in zip(groups[0][:4], "GRIZ"))
    assert len(F.find_tile_bands([str(tmp_path)], bands, min_bands=1, use_calexp=True)) == 2
    assert len(F.find_tile_bands([str(tmp_path)], bands, min_bands=1, use_calexp=False)) == 1
    tile, p2r = F.load_tile_bands(groups[0])
    assert tile.shape == (5, 96, 128) and tile.dtype == np.float32 and np.isnan(tile[4]).all()
    for i, b in enumerate("GRIZ"):
        assert np.array_equal(tile[i], tiles[b])
    ra, dec = p2r([100.0], [80.0])
    assert abs(ra[0] - 150.25) < 1e-9 and abs(dec[0] - 2.5) < 1e-9


def test_tile_loader_over_fits_files(tmp_path):
    """FITS files -> the nested per-tile batches of the tile loader (host mode), ra/dec from the header's WCS, argument
    order as the reference passes it (row centres as x, utils/dataloaders.py:528-533)."""
    from sky_embeddings_b200 import ingest
    rng = np.random.default_rng(4)
    bands = ["G", "R", "I", "Z", "Y"]
    planes = []
    for b in bands:
        img = (rng.standard_normal((200, 173)) * 4).astype(np.float32)
        planes.append(img)
        F.write_image(str(tmp_path / f"calexp-HSC-{b}-1-0,0.fits"), img, HDR)
    groups = F.find_tile_bands([str(tmp_path)], bands, 5)
    tile, p2r = F.load_tile_bands(groups[0])
    (cut, masks, rd), = list(ingest.TileLoader([(tile, p2r)], batch_size=4, img_size=64, overlap=0.4))
    coords = ingest.generate_overlap_coords((200, 173), 64, 0.4)
    assert cut.shape == (1, len(coords) // 4, 4, 5, 64, 64)
    h0, w0 = coords[7]
    want = np.stack(planes)[:, h0:h0 + 64, w0:w0 + 64].copy()
    want[want < -3] = -3
    assert np.array_equal(cut.reshape(-1, 5, 64, 64)[7].numpy(), want)
    ra, dec = F.TanWcs(HDR).all_pix2world([h0 + 32], [w0 + 32], 0)
    assert np.allclose(rd.reshape(-1, 2)[7].numpy(), [ra[0], dec[0]], atol=1e-4)

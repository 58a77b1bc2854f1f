"""CPU tests of the ingest layer: the HDF5 subset reader / writer, the overlap-coordinate list against the
reference's generate_overlap_coords (golden, utils/dataloaders.py:481-509), and the host mirrors of the loaders."""
import hashlib
import os

import numpy as np
import pytest

from sky_embeddings_b200 import h5lite, ingest
from tests import golden_inputs as G


def _bank_arrays(n=11, C=5, S=64, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, C, S, S)).astype(np.float32) * 3
    x[2, 1] = np.nan
    return dict(cutouts=x, ra=rng.random(n).astype("f") * 360, dec=(rng.random(n).astype("f") - 0.5) * 90,
                zspec=rng.random(n).astype("f"), zspec_err=rng.random(n).astype("f"))


@pytest.mark.parametrize("userblock", [0, 512, 2048])
def test_h5lite_round_trip(tmp_path, userblock):
    ds = _bank_arrays()
    ds["ids"] = np.arange(11, dtype=np.int64)
    ds["be"] = np.arange(6, dtype=">f8").reshape(2, 3)
    p = h5lite.write_h5(str(tmp_path / "bank.h5"), ds, userblock=userblock)
    with h5lite.H5File(p) as f:
        assert sorted(f.keys()) == sorted(ds)
        for k, v in ds.items():
            got = np.asarray(f[k])
            assert got.shape == v.shape and got.dtype.kind == v.dtype.kind and got.dtype.itemsize == v.dtype.itemsize
            assert np.array_equal(got, v, equal_nan=True), k
        assert len(f["cutouts"]) == 11
        with pytest.raises(KeyError):
            f["nope"]


def test_h5lite_reads_a_libhdf5_file():
    """A file written by libhdf5 itself (MATLAB v7.3 = HDF5 with a 512-byte user block, old-style group, layout v2
    contiguous float64) ships with scipy's test data; skip where scipy has no tests installed."""
    import scipy.io
    p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(p):
        pytest.skip("scipy test data not installed")
    with h5lite.H5File(p) as f:
        assert f.keys() == ["testdouble"]
        a = np.asarray(f["testdouble"])
        assert a.shape == (9, 1) and np.allclose(a[:, 0], np.arange(9) * np.pi / 4)


def test_h5lite_rejects_garbage_and_truncation(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(h5lite.H5Error):
        h5lite.H5File(str(p))
    good = h5lite.write_h5(str(tmp_path / "good.h5"), _bank_arrays(n=3))
    data = open(good, "rb").read()
    (tmp_path / "cut.h5").write_bytes(data[: len(data) // 2])
    with pytest.raises(h5lite.H5Error):
        with h5lite.H5File(str(tmp_path / "cut.h5")) as f:
            np.asarray(f["cutouts"]).sum()


def test_h5lite_matches_h5py_when_available(tmp_path):
    h5py = pytest.importorskip("h5py")
    ds = _bank_arrays()
    p = str(tmp_path / "by_h5py.h5")
    with h5py.File(p, "w") as f:                    # exactly data_processing/utils.py:346-350
        for k, v in ds.items():
            f.create_dataset(k, v.shape, dtype="f")[...] = v
    with h5lite.H5File(p) as f:
        for k, v in ds.items():
            assert np.array_equal(np.asarray(f[k]), v, equal_nan=True)
    q = h5lite.write_h5(str(tmp_path / "by_lite.h5"), ds)
    with h5py.File(q, "r") as f:
        for k, v in ds.items():
            assert np.array_equal(f[k][...], v, equal_nan=True)


def test_overlap_coords_match_reference_golden():
    g = G.load("ingest")
    for i, (H, W, size, ov) in enumerate(g["coord_cases"]):
        got = ingest.generate_overlap_coords((int(H), int(W)), int(size), float(ov))
        assert got.dtype == np.int32 and np.array_equal(got, g[f"coords.{i}"]), (H, W, size, ov)
    # the reference's edge test is on H % step: 112 with step 48 pins an extra row on top of a regular one
    c = ingest.generate_overlap_coords((112, 112), 64, 0.25)
    assert len(c) == 9 and len({tuple(r) for r in c.tolist()}) < 9


def test_h5_loader_host_mode_follows_the_reference_items(tmp_path):
    """device=None: the batches H5Dataset + DataLoader(shuffle=False) produce (utils/dataloaders.py:284-329)."""
    import torch
    rng = np.random.default_rng(5)
    n, C, S = 23, 5, 96                               # stored larger than img_size -> central crop
    x = (rng.standard_normal((n, C, S, S)) * 4).astype(np.float32)
    x[4, 2] = np.nan
    ra, dec = rng.random(n).astype("f"), rng.random(n).astype("f")
    p = h5lite.write_h5(str(tmp_path / "bank.h5"), dict(cutouts=x, ra=ra, dec=dec))
    src = ingest.H5Cutouts(p, img_size=64, pixel_min=-3.0)
    idx = np.array([1, 2, 5, 7, 8, 13, 21, 22, 4])
    ld = src.loader(indices=idx, batch_size=4)
    assert len(ld) == 3
    got_x, got_rd = [], []
    for cut, mask, rd in ld:
        assert cut.dtype == torch.float32 and mask.shape == cut.shape and float(mask.abs().sum()) == 0.0
        got_x.append(cut.numpy()); got_rd.append(rd.numpy())
    want = x[idx].copy()
    want[want < -3.0] = -3.0
    want = want[:, :, 48 - 32:48 + 32, 48 - 32:48 + 32]
    assert np.array_equal(np.concatenate(got_x), want, equal_nan=True)
    assert np.array_equal(np.concatenate(got_rd), np.stack((ra[idx], dec[idx]), -1))
    src.close()


def test_tile_loader_host_mode_matches_reference_golden():
    g = G.load("ingest")
    _, tile = G.ingest_inputs(g)
    ld = ingest.TileLoader([(tile, G.tile_pix_to_radec)], batch_size=4, img_size=64, overlap=0.4, pixel_min=-3.0)
    (cut, masks, rd), = list(ld)
    n = len(g["coords.0"])
    M = n // 4
    assert cut.shape == (1, M, 4, 3, 64, 64) and masks.shape == (1, M, 4) and rd.shape == (1, M, 4, 2)
    flat = cut.reshape(-1, 3, 64, 64).numpy()
    assert np.array_equal(flat[:2], g["tile_cutouts_first"], equal_nan=True)
    assert M * 4 == n, "the case was chosen so that no cutout is dropped"
    assert hashlib.sha256(np.ascontiguousarray(flat).tobytes()).hexdigest() == str(g["tile_cutouts_sha"])
    assert np.allclose(rd.reshape(-1, 2).numpy(), g["tile_ra_dec"], rtol=0, atol=1e-5)

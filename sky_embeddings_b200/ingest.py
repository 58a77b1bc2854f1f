"""Bank ingest formats in front of the resident bank (a10, f2, f3).

What the reference does on the way from disk to ``mae_simsearch``:

  * h5 bank (``similarity_search.py``): ``h5_snr`` over the whole file (utils/misc.py:165-180, 5000 cutouts at a time
    through numpy), ``np.nanmin`` over the first five channels and the ``snr_range`` window give ``test_indices``
    (similarity_search.py:124-130); ``build_h5_dataloader(..., shuffle=False, indices=test_indices)`` then opens the
    file once PER ITEM (utils/dataloaders.py:289), clips at -3, crops the centre and yields
    ``(cutout, zeros_like(cutout), [ra, dec])`` batches.
  * FITS tiles (``sky_sim_search.py``): one tile per dataset item; ``generate_overlap_coords`` +
    ``overlapping_cutouts`` (utils/dataloaders.py:481-536) cut it into overlapping 64 x 64 cutouts on the host,
    clipped and reshaped to ``[M, batch, C, 64, 64]`` nested batches (:657-679).

Here the file is memory-mapped once (``h5py`` when importable, else ``h5lite``), staged to the GPU in large pinned
chunks, and the per-pixel work runs in CUDA (``csrc/pixel_prep.cu``): the S/N statistic, clipping / centre crop, and
the tile -> cutout gather.  Index selection and coordinate lists stay on the host (they are index arithmetic).
No step here has a CPU fallback for the pixel work: without the CUDA library these functions raise.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import _lib
from . import h5lite


def open_h5(path):
    """``h5py.File`` when h5py is importable, else the pure-Python reader (contiguous datasets only)."""
    try:
        import h5py  # noqa: PLC0415
    except ImportError:
        h5py = None
    if h5py is not None and hasattr(h5py, "File"):
        return h5py.File(path, "r")
    return h5lite.H5File(path)


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _nan_or(v):
    return float("nan") if v is None else float(v)


# ---------------------------------------------------------------------------------------------------------------
# device kernels
# ---------------------------------------------------------------------------------------------------------------
def snr_device(cutouts, n_central_pix=8, n_min_channels=5):
    """calculate_snr (utils/misc.py:119-163) on the GPU.  cutouts: CUDA f32 [n, C, H, H].
    Returns (snr [n, C], min_snr [n] = nanmin over the first n_min_channels channels)."""
    if not cutouts.is_cuda:
        raise RuntimeError("snr_device needs a CUDA tensor: the S/N statistic has no CPU fallback here")
    x = cutouts.contiguous().float()
    n, C, H, W = x.shape
    snr = torch.empty((n, C), dtype=torch.float32, device=x.device)
    mn = torch.empty((n,), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    _lib.check(lib.sky_pixel_snr(x.data_ptr(), n, C, H, W, int(n_central_pix), int(n_min_channels), snr.data_ptr(),
                                 mn.data_ptr(), x.device.index or 0, _stream(x.device)))
    return snr, mn


def center_clip_device(src, img_size, pixel_min=-3.0, pixel_max=None):
    """The per-item steps of H5Dataset.__getitem__ (utils/dataloaders.py:291-300) for a whole batch on the GPU."""
    if not src.is_cuda:
        raise RuntimeError("center_clip_device needs a CUDA tensor")
    x = src.contiguous().float()
    n, C, Hs, Ws = x.shape
    size = min(img_size, Hs, Ws) if (Hs > img_size or Ws > img_size) else Hs
    if Hs != Ws and not (Hs > img_size or Ws > img_size):
        raise ValueError("non-square cutouts are only supported when they are cropped to img_size")
    out = torch.empty((n, C, size, size), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    _lib.check(lib.sky_center_clip(x.data_ptr(), n, C, Hs, Ws, size, _nan_or(pixel_min), _nan_or(pixel_max), out.data_ptr(),
                                   x.device.index or 0, _stream(x.device)))
    return out


def generate_overlap_coords(img_shape, cutout_size, overlap):
    """Top-left (h, w) of every overlapping cutout of a tile, in the reference's order INCLUDING its edge rules
    (utils/dataloaders.py:481-509): a regular grid of step int(size * (1 - overlap)), then one extra row pinned to the
    bottom edge when H % step != 0, one extra column pinned to the right edge when W % step != 0, and the corner when
    both hold.  (The edge tests are on H % step, not on (H - size) % step, so an extra row can duplicate a regular
    one; that is reproduced.)  Returns int32 [n, 2]."""
    H, W = int(img_shape[0]), int(img_shape[1])
    step = int(cutout_size * (1 - overlap))
    if step < 1:
        raise ValueError("overlap leaves no step between cutouts")
    hs = np.arange(0, H - cutout_size + 1, step, dtype=np.int32)
    ws = np.arange(0, W - cutout_size + 1, step, dtype=np.int32)
    parts = [np.stack(np.meshgrid(hs, ws, indexing="ij"), -1).reshape(-1, 2)]
    if H % step != 0:
        parts.append(np.stack((np.full_like(ws, H - cutout_size), ws), -1))
    if W % step != 0:
        parts.append(np.stack((hs, np.full_like(hs, W - cutout_size)), -1))
    if H % step != 0 and W % step != 0:
        parts.append(np.array([[H - cutout_size, W - cutout_size]], dtype=np.int32))
    return np.concatenate(parts).astype(np.int32)


def tile_cutouts_device(tile, coords, img_size, pixel_min=-3.0, pixel_max=None):
    """overlapping_cutouts + clipping (utils/dataloaders.py:511-536, :657-661) as one gather kernel.
    tile: CUDA f32 [C, H, W]; coords: int32 [n, 2] (host array or CUDA tensor) -> CUDA f32 [n, C, img_size, img_size]."""
    if not tile.is_cuda:
        raise RuntimeError("tile_cutouts_device needs the tile on the GPU")
    t = tile.contiguous().float()
    C, H, W = t.shape
    c = torch.as_tensor(coords, dtype=torch.int32).to(t.device).contiguous()
    n = c.shape[0]
    if n and (int(c[:, 0].max()) + img_size > H or int(c[:, 1].max()) + img_size > W or int(c.min()) < 0):
        raise ValueError("cutout coordinates reach outside the tile")
    out = torch.empty((n, C, img_size, img_size), dtype=torch.float32, device=t.device)
    lib = _lib.load()
    _lib.check(lib.sky_tile_cutouts(t.data_ptr(), C, H, W, c.data_ptr(), n, img_size, _nan_or(pixel_min), _nan_or(pixel_max),
                                    out.data_ptr(), t.device.index or 0, _stream(t.device)))
    return out


# ---------------------------------------------------------------------------------------------------------------
# h5 bank file
# ---------------------------------------------------------------------------------------------------------------
class H5Cutouts:
    """The reference's bank file (data_processing/utils.py:346-350): ``cutouts`` f32 [N, C, H, W], ``ra``, ``dec`` f32 [N]."""

    def __init__(self, path, img_size=64, pixel_min=-3.0, pixel_max=None):
        self.path = path
        self.f = open_h5(path)
        self.cutouts = self.f["cutouts"]
        self.ra = np.asarray(self.f["ra"][:], dtype=np.float32)
        self.dec = np.asarray(self.f["dec"][:], dtype=np.float32)
        self.img_size, self.pixel_min, self.pixel_max = img_size, pixel_min, pixel_max
        if self.cutouts.ndim != 4:
            raise ValueError(f"{path}: 'cutouts' must be [N, C, H, W], found shape {tuple(self.cutouts.shape)}")

    def __len__(self):
        return int(self.cutouts.shape[0])

    def close(self):
        self.f.close()

    def _stage(self, rows, device):
        """Rows of the file -> CUDA f32 [n, C, H, W]: one gathered host copy into pinned memory, one async H2D."""
        if isinstance(rows, slice):
            host = np.array(self.cutouts[rows], dtype=np.float32)          # a copy: the map is read-only
        else:
            rows = np.asarray(rows)
            host = np.array(self.cutouts[rows] if len(rows) else
                            np.zeros((0,) + tuple(self.cutouts.shape[1:]), np.float32), dtype=np.float32)
        t = torch.from_numpy(host)
        if device.type == "cuda" and t.numel():
            t = t.pin_memory()
        return t.to(device, non_blocking=True)

    def snr(self, device, n_central_pix=8, batch_size=5000, num_samples=None):
        """h5_snr (utils/misc.py:165-180): S/N of every cutout and channel, [N, C] f32 on `device` (and the nanmin over
        the first five channels, [N]); the file is streamed in `batch_size` slices."""
        device = torch.device(device)
        n = len(self) if num_samples is None else min(int(num_samples), len(self))
        C = self.cutouts.shape[1]
        snr = torch.empty((n, C), dtype=torch.float32, device=device)
        mn = torch.empty((n,), dtype=torch.float32, device=device)
        for i in range(0, n, batch_size):
            e = min(n, i + batch_size)
            s, m = snr_device(self._stage(slice(i, e), device), n_central_pix, 5)
            snr[i:e], mn[i:e] = s, m
        return snr, mn

    def select_snr(self, device, snr_range, n_central_pix=8, batch_size=5000):
        """test_indices of similarity_search.py:124-130: ascending rows whose nanmin S/N over the first five channels
        lies strictly inside snr_range.  Returns (int64 numpy indices, min_snr tensor)."""
        _, mn = self.snr(device, n_central_pix, batch_size)
        keep = (mn > float(snr_range[0])) & (mn < float(snr_range[1]))        # NaN compares false, as in numpy
        return torch.nonzero(keep).flatten().cpu().numpy().astype(np.int64), mn

    def loader(self, indices=None, batch_size=64, device=None):
        return H5CutoutLoader(self, indices, batch_size, device)

    def pixel_bank(self, device, indices=None, chunk_items=4096):
        """Pixel-space bank (BASELINE config 5) of the selected rows, clipped / cropped like the loader's items."""
        from .engine import PixelBank  # noqa: PLC0415
        device = torch.device(device)
        idx = np.arange(len(self)) if indices is None else np.asarray(indices)
        C = self.cutouts.shape[1]
        size = min(self.img_size, self.cutouts.shape[2], self.cutouts.shape[3])
        pb = PixelBank(len(idx), C, size, size, device=device)
        for i in range(0, len(idx), chunk_items):
            rows = idx[i:i + chunk_items]
            pb.upload(center_clip_device(self._stage(rows, device), self.img_size, self.pixel_min, self.pixel_max), item0=i)
        return pb


class H5CutoutLoader:
    """Drop-in for ``build_h5_dataloader(fn, batch_size, ..., max_mask_ratio=None, shuffle=False, indices=indices)``
    (utils/dataloaders.py:134-153) on the search path: iterable with ``len()`` yielding
    ``(cutouts [B, C, S, S], masks = zeros_like(cutouts), ra_dec [B, 2])`` in the order of ``indices``.
    device=None yields pinned CPU tensors, exactly what the reference loader yields (mae_simsearch moves them);
    with a CUDA device the batch is clipped / cropped on the GPU and stays there."""

    def __init__(self, src, indices=None, batch_size=64, device=None):
        self.src = src
        self.indices = np.arange(len(src), dtype=np.int64) if indices is None else np.asarray(indices, dtype=np.int64)
        self.batch_size = int(batch_size)
        self.device = None if device is None else torch.device(device)

    def __len__(self):
        return math.ceil(len(self.indices) / self.batch_size)

    def __iter__(self):
        s = self.src
        for i in range(0, len(self.indices), self.batch_size):
            rows = self.indices[i:i + self.batch_size]
            ra_dec = torch.from_numpy(np.stack((s.ra[rows], s.dec[rows]), -1).astype(np.float32))
            if self.device is not None and self.device.type == "cuda":
                x = center_clip_device(s._stage(rows, self.device), s.img_size, s.pixel_min, s.pixel_max)
                yield x, torch.zeros_like(x), ra_dec.to(self.device)
                continue
            x = np.array(s.cutouts[rows], dtype=np.float32)                      # host mirror of :291-300
            if s.pixel_min is not None:
                x[x < s.pixel_min] = s.pixel_min
            if s.pixel_max is not None:
                x[x > s.pixel_max] = s.pixel_max
            if x.shape[2] > s.img_size or x.shape[3] > s.img_size:
                r0, c0 = x.shape[2] // 2 - s.img_size // 2, x.shape[3] // 2 - s.img_size // 2
                x = np.ascontiguousarray(x[:, :, r0:r0 + s.img_size, c0:c0 + s.img_size])
            t = torch.from_numpy(x)
            yield t, torch.zeros_like(t), ra_dec


# ---------------------------------------------------------------------------------------------------------------
# FITS-tile streaming (the other caller, sky_sim_search.py:137-164)
# ---------------------------------------------------------------------------------------------------------------
class TileLoader:
    """Drop-in for ``build_fits_dataloader(..., use_overlap=True, ra_dec=True, shuffle=False)`` on the search path,
    over tiles that are already arrays (astropy is not in this image: the caller -- or ``load_tile_npy`` -- supplies
    ``(tile [C, H, W] f32 with NaN for missing bands, pix_to_radec or None)``).

    Each item is one tile as NESTED batches, like FitsDataset.__getitem__ (utils/dataloaders.py:641-679) seen through
    the DataLoader's batch_size=1 collation (:131-133): ``(cutouts [1, M, B, C, S, S], masks [1, M, B], ra_dec
    [1, M, B, 2])``, M = n_cutouts // B, the remainder dropped.  With a CUDA device the tile is uploaded once and cut
    up by the gather kernel; the next tile's upload is issued on a side stream while the caller works on this one
    (double buffering)."""

    def __init__(self, tiles, batch_size=64, img_size=64, overlap=0.4, pixel_min=-3.0, pixel_max=None, device=None):
        self.tiles = tiles
        self.batch_size, self.img_size, self.overlap = int(batch_size), int(img_size), float(overlap)
        self.pixel_min, self.pixel_max = pixel_min, pixel_max
        self.device = None if device is None else torch.device(device)

    def __len__(self):
        return len(self.tiles)

    def _radec(self, pix_to_radec, coords):
        if pix_to_radec is None:
            return np.zeros((len(coords), 2), np.float32)
        hc = [int(h) + self.img_size // 2 for h, _ in coords]
        wc = [int(w) + self.img_size // 2 for _, w in coords]
        ra, dec = pix_to_radec(hc, wc)                                          # argument order as :530-533
        return np.vstack((ra, dec)).T.astype(np.float32)

    def _fetch(self, i, stream=None):
        tile, p2r = self.tiles[i]
        tile = np.asarray(tile, dtype=np.float32)
        coords = generate_overlap_coords(tile.shape[1:], self.img_size, self.overlap)
        ra_dec = self._radec(p2r, coords)
        if self.device is not None and self.device.type == "cuda":
            host = torch.from_numpy(np.ascontiguousarray(tile)).pin_memory()
            with torch.cuda.stream(stream):
                dev_tile = host.to(self.device, non_blocking=True)
                dev_coords = torch.from_numpy(coords).pin_memory().to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            return dev_tile, dev_coords, ra_dec, ev, host
        return tile, coords, ra_dec, None, None

    def __iter__(self):
        B, S = self.batch_size, self.img_size
        cuda = self.device is not None and self.device.type == "cuda"
        side = torch.cuda.Stream(self.device) if cuda else None
        nxt = self._fetch(0, side) if len(self.tiles) else None
        for i in range(len(self.tiles)):
            tile, coords, ra_dec, ev, _keep = nxt
            nxt = self._fetch(i + 1, side) if i + 1 < len(self.tiles) else None      # overlaps this tile's work
            M = len(coords) // B
            if cuda:
                torch.cuda.current_stream(self.device).wait_event(ev)
                tile.record_stream(torch.cuda.current_stream(self.device))
                coords.record_stream(torch.cuda.current_stream(self.device))
                cut = tile_cutouts_device(tile, coords[:M * B], S, self.pixel_min, self.pixel_max)
                rd = torch.from_numpy(ra_dec[:M * B]).to(self.device)
            else:
                cut = np.stack([tile[:, h:h + S, w:w + S] for h, w in coords[:M * B]]) if M else np.zeros((0, tile.shape[0], S, S), np.float32)
                if self.pixel_min is not None:
                    cut[cut < self.pixel_min] = self.pixel_min
                if self.pixel_max is not None:
                    cut[cut > self.pixel_max] = self.pixel_max
                cut = torch.from_numpy(cut.astype(np.float32))
                rd = torch.from_numpy(ra_dec[:M * B])
            C = cut.shape[1]
            yield (cut.reshape(1, M, B, C, S, S), torch.zeros((1, M, B), device=cut.device), rd.reshape(1, M, B, 2))


def load_tile_npy(path):
    """A tile saved as ``.npy`` [C, H, W] f32 (NaN planes for missing bands, as load_fits_bands builds them,
    utils/dataloaders.py:430-437), memory-mapped; no WCS (ra/dec come back as zeros)."""
    return np.load(path, mmap_mode="r"), None

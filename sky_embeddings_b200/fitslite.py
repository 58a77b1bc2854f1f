"""fitslite -- a small FITS image reader with a TAN(-SIP) pixel -> sky transform, for the tile-streaming caller.

The reference reads its sky tiles with astropy (`fits.open(fn)[1].data`, `WCS(hdul[1].header).all_pix2world`,
/root/reference/utils/dataloaders.py:404-424) and finds the per-band files of a patch by name
(`find_HSC_bands`, :330-379).  astropy is not in this image; what that code needs is small:

  * FITS structure (FITS Standard 4.0): 2880-byte blocks, 80-character header cards up to END, then the data of
    |BITPIX| / 8 * NAXIS1 * ... bytes (big-endian), padded to a block; HDU 1 = the first extension.  Uncompressed
    IMAGE HDUs of BITPIX -32 / -64 / 8 / 16 / 32 / 64 with BSCALE / BZERO; tile-compressed images (ZIMAGE) raise.
    The pixel array is memory-mapped: nothing is copied until a tile is staged for the GPU.
  * the gnomonic (TAN) projection with a CD (or PC + CDELT) matrix and optional SIP distortion polynomials
    (Calabretta & Greisen 2002, eq. 2 and 54-55; Shupe et al. 2005 for SIP): what `all_pix2world(x, y, 0)` evaluates
    for HSC coadd headers (`RA---TAN`, `DEC--TAN`, `RA---TAN-SIP`).  **Parity unpinned**: there is no astropy here to
    pin it against; tests check it against the defining identities (reference pixel -> CRVAL, great-circle distance
    = arctan of the tangent-plane radius, an independent inverse) only.

`find_tile_bands` / `load_tile_bands` follow the reference's file-name convention and missing-band handling.
"""
from __future__ import annotations

import glob
import math
import os

import numpy as np

BLOCK = 2880


class FitsError(ValueError):
    pass


def _parse_card(card):
    key = card[:8].strip()
    if card[8:10] != "= " or key in ("COMMENT", "HISTORY", ""):
        return key, None
    val = card[10:]
    if val.lstrip().startswith("'"):                      # string: up to the closing quote ('' = a quote)
        s = val.lstrip()[1:]
        out, i = [], 0
        while i < len(s):
            if s[i] == "'":
                if i + 1 < len(s) and s[i + 1] == "'":
                    out.append("'"); i += 2
                    continue
                break
            out.append(s[i]); i += 1
        return key, "".join(out).rstrip()
    val = val.split("/")[0].strip()
    if val in ("T", "F"):
        return key, val == "T"
    try:
        return key, int(val)
    except ValueError:
        try:
            return key, float(val.replace("D", "E"))
        except ValueError:
            return key, val


def read_hdus(path):
    """[(header dict, data offset, data bytes)] of every HDU of the file."""
    size = os.path.getsize(path)
    hdus, off = [], 0
    with open(path, "rb") as f:
        while off < size:
            hdr, done = {}, False
            while not done:
                f.seek(off)
                block = f.read(BLOCK)
                if len(block) < BLOCK:
                    if not hdr and not hdus:
                        raise FitsError(f"{path}: not a FITS file (short header block)")
                    return hdus
                off += BLOCK
                for i in range(0, BLOCK, 80):
                    card = block[i:i + 80].decode("ascii", "replace")
                    if card.startswith("END") and card[3:].strip() == "":
                        done = True
                        break
                    k, v = _parse_card(card)
                    if v is not None and k not in hdr:
                        hdr[k] = v
            if not hdus and hdr.get("SIMPLE") is not True:
                raise FitsError(f"{path}: not a FITS file (no SIMPLE = T)")
            naxis = int(hdr.get("NAXIS", 0))
            n = 1
            for a in range(1, naxis + 1):
                n *= int(hdr.get(f"NAXIS{a}", 0))
            if naxis == 0:
                n = 0
            nbytes = abs(int(hdr.get("BITPIX", 8))) // 8 * int(hdr.get("GCOUNT", 1)) * (int(hdr.get("PCOUNT", 0)) + n)
            hdus.append((hdr, off, nbytes))
            off += (nbytes + BLOCK - 1) // BLOCK * BLOCK
    return hdus


_DTYPES = {-32: ">f4", -64: ">f8", 8: "u1", 16: ">i2", 32: ">i4", 64: ">i8"}


def read_image(path, hdu=1):
    """(data [NAXIS2, NAXIS1] memory-mapped (big-endian as stored; scaled copy when BSCALE / BZERO say so), header)."""
    hdus = read_hdus(path)
    if hdu >= len(hdus):
        raise FitsError(f"{path}: has {len(hdus)} HDUs, HDU {hdu} requested")
    hdr, off, nbytes = hdus[hdu]
    if hdr.get("ZIMAGE") is True or hdr.get("XTENSION", "IMAGE").strip() == "BINTABLE":
        raise FitsError(f"{path}: HDU {hdu} is a tile-compressed image / table; decompress it (funpack) or read it with astropy")
    if int(hdr.get("NAXIS", 0)) != 2:
        raise FitsError(f"{path}: HDU {hdu} is not a 2-D image (NAXIS = {hdr.get('NAXIS')})")
    bitpix = int(hdr["BITPIX"])
    if bitpix not in _DTYPES:
        raise FitsError(f"{path}: BITPIX {bitpix}")
    shape = (int(hdr["NAXIS2"]), int(hdr["NAXIS1"]))
    if off + nbytes > os.path.getsize(path):
        raise FitsError(f"{path}: data of HDU {hdu} runs past the end of the file (truncated?)")
    data = np.memmap(path, dtype=np.dtype(_DTYPES[bitpix]), mode="r", offset=off, shape=shape)
    bscale, bzero = float(hdr.get("BSCALE", 1.0)), float(hdr.get("BZERO", 0.0))
    if bscale != 1.0 or bzero != 0.0:
        data = data.astype(np.float64) * bscale + bzero
    return data, hdr


class TanWcs:
    """pixel -> (ra, dec) in degrees for `RA---TAN[-SIP]` / `DEC--TAN[-SIP]` headers."""

    def __init__(self, hdr):
        c1, c2 = str(hdr.get("CTYPE1", "")), str(hdr.get("CTYPE2", ""))
        if not (c1.startswith("RA---TAN") and c2.startswith("DEC--TAN")):
            raise FitsError(f"unsupported projection CTYPE1={c1!r} CTYPE2={c2!r}: only RA---TAN / DEC--TAN (optionally -SIP)")
        self.crpix = (float(hdr["CRPIX1"]), float(hdr["CRPIX2"]))
        self.crval = (float(hdr["CRVAL1"]), float(hdr["CRVAL2"]))
        if "CD1_1" in hdr:
            self.cd = np.array([[hdr.get("CD1_1", 0.0), hdr.get("CD1_2", 0.0)], [hdr.get("CD2_1", 0.0), hdr.get("CD2_2", 0.0)]], float)
        else:
            pc = np.array([[hdr.get("PC1_1", 1.0), hdr.get("PC1_2", 0.0)], [hdr.get("PC2_1", 0.0), hdr.get("PC2_2", 1.0)]], float)
            self.cd = np.diag([float(hdr.get("CDELT1", 1.0)), float(hdr.get("CDELT2", 1.0))]) @ pc
        self.lonpole = float(hdr.get("LONPOLE", 180.0))
        self.sip = None
        if c1.endswith("-SIP"):
            def poly(prefix):
                order = int(hdr.get(f"{prefix}_ORDER", 0))
                return [(p, q, float(hdr[f"{prefix}_{p}_{q}"])) for p in range(order + 1) for q in range(order + 1 - p)
                        if f"{prefix}_{p}_{q}" in hdr]
            self.sip = (poly("A"), poly("B"))

    def all_pix2world(self, x, y, origin=0):
        """As astropy's method: x = FITS axis 1 (column), y = axis 2 (row), `origin` 0 for 0-based pixels."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        u = x + (1 - origin) - self.crpix[0]
        v = y + (1 - origin) - self.crpix[1]
        if self.sip is not None:
            du = sum(c * u ** p * v ** q for p, q, c in self.sip[0])
            dv = sum(c * u ** p * v ** q for p, q, c in self.sip[1])
            u, v = u + du, v + dv
        xi = np.deg2rad(self.cd[0, 0] * u + self.cd[0, 1] * v)
        eta = np.deg2rad(self.cd[1, 0] * u + self.cd[1, 1] * v)
        r = np.hypot(xi, eta)
        phi = np.arctan2(xi, -eta)
        theta = np.arctan2(1.0, r)                          # tan(theta) = 1 / r  (r in radians)
        a0, d0, phip = np.deg2rad(self.crval[0]), np.deg2rad(self.crval[1]), np.deg2rad(self.lonpole)
        dphi = phi - phip
        sin_d = np.sin(theta) * np.sin(d0) + np.cos(theta) * np.cos(d0) * np.cos(dphi)
        dec = np.arcsin(np.clip(sin_d, -1.0, 1.0))
        ra = a0 + np.arctan2(-np.cos(theta) * np.sin(dphi), np.sin(theta) * np.cos(d0) - np.cos(theta) * np.sin(d0) * np.cos(dphi))
        return np.rad2deg(ra) % 360.0, np.rad2deg(dec)


def find_tile_bands(fits_paths, bands, min_bands=2, use_calexp=True, verbose=0):
    """The reference's find_HSC_bands (utils/dataloaders.py:330-379): `[calexp-]HSC-<band>-<tract>-<patch>.fits` files
    grouped by patch; a missing band is the string 'None'; patches with fewer than min_bands bands are dropped."""
    patch_files = {}
    for d in fits_paths:
        for file_path in glob.glob(f"{d}/*.fits"):
            name = file_path.split("/")[-1]
            if (use_calexp and name.startswith("calexp-")) or (not use_calexp and not name.startswith("calexp-")):
                parts = name.split("-")
                if len(parts) < 3:
                    continue
                band, patch = parts[-3], "-".join(parts[-2:])
                if band in bands:
                    patch_files.setdefault(patch, {b: "None" for b in bands})[band] = file_path
    out = []
    for patch, avail in patch_files.items():
        cur = [avail[b] for b in bands]
        if len([f for f in cur if f != "None"]) >= min_bands:
            out.append(cur)
    if verbose:
        print(f"Found {len(out)} patches with at least {min_bands} of the {bands} bands.")
    return out


def load_tile_bands(patch_filenames, return_wc=True):
    """The reference's load_fits_bands (utils/dataloaders.py:381-437): (tile [C, H, W] with NaN planes for missing or
    unreadable bands, pix_to_radec or None).  pix_to_radec(x, y) = all_pix2world(x, y, 0) of the FIRST readable band,
    called by the tile loader with (row centres, column centres) exactly as the reference calls it (:528-533)."""
    imgs, ref_shape, wcs = [], None, None
    for fn in patch_filenames:
        if fn == "None":
            imgs.append(None)
            continue
        try:
            data, hdr = read_image(fn, 1)
            if ref_shape is None:
                ref_shape = data.shape
            imgs.append(data)
            if wcs is None and return_wc:
                wcs = TanWcs(hdr)
        except Exception as e:       # noqa: BLE001 -- the reference prints and fills the band with NaN
            print(f"Error opening {fn}: {e}")
            imgs.append(None)
    if ref_shape is None:
        raise FitsError(f"no readable band among {patch_filenames}")
    tile = np.stack([np.full(ref_shape, np.nan, np.float32) if im is None else np.asarray(im, dtype=np.float32) for im in imgs])
    return tile, (None if wcs is None else (lambda x, y: wcs.all_pix2world(x, y, 0)))


def write_image(path, data, header=None, primary_empty=True):
    """Write a 2-D float32 image as HDU 1 (after an empty primary HDU, like the HSC calexp files) -- fixtures / exports."""
    data = np.asarray(data, dtype=">f4")

    def cards(items):
        out = b""
        for k, v in items:
            if isinstance(v, bool):
                s = f"{k:<8}= {'T' if v else 'F':>20}"
            elif isinstance(v, (int, np.integer)):
                s = f"{k:<8}= {int(v):>20}"
            elif isinstance(v, (float, np.floating)):
                s = f"{k:<8}= {float(v):>20.13E}"
            else:
                s = f"{k:<8}= '{str(v):<8}'"
            out += s.ljust(80).encode("ascii")
        out += b"END".ljust(80)
        return out.ljust((len(out) + BLOCK - 1) // BLOCK * BLOCK, b" ")

    with open(path, "wb") as f:
        if primary_empty:
            f.write(cards([("SIMPLE", True), ("BITPIX", 8), ("NAXIS", 0), ("EXTEND", True)]))
            head = [("XTENSION", "IMAGE"), ("BITPIX", -32), ("NAXIS", 2), ("NAXIS1", data.shape[1]), ("NAXIS2", data.shape[0]),
                    ("PCOUNT", 0), ("GCOUNT", 1)]
        else:
            head = [("SIMPLE", True), ("BITPIX", -32), ("NAXIS", 2), ("NAXIS1", data.shape[1]), ("NAXIS2", data.shape[0])]
        f.write(cards(head + list((header or {}).items())))
        raw = data.tobytes()
        f.write(raw.ljust((len(raw) + BLOCK - 1) // BLOCK * BLOCK, b"\0"))
    return path

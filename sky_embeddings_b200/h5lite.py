"""h5lite -- a small pure-Python reader (and fixture writer) for the HDF5 subset the reference's bank files use.

The reference stores a bank as an HDF5 file with contiguous, uncompressed, little-endian datasets created by
``f.create_dataset(name, shape, dtype='f')`` (/root/reference/data_processing/utils.py:346-350: ``cutouts``
[N, C, 64, 64], ``ra``, ``dec``, ``zspec``, ``zspec_err`` [N]) and reads them item by item through ``h5py``
(/root/reference/utils/dataloaders.py:289-304).  ``h5py`` / libhdf5 are not in this image, and nothing more than
"find the dataset, map its bytes" is needed to feed a device-resident bank, so this module parses the file format
directly (HDF5 File Format Specification, version 3.0):

  * superblock versions 0-3 (a user block in front is honoured: base address);
  * groups: old style (symbol table message -> v1 B-tree -> symbol nodes -> local heap; what ``h5py`` writes by
    default) and compact new style (link messages in a version-2 object header; ``libver='latest'``);
  * object headers version 1 and 2, continuation blocks;
  * dataspace v1 / v2 (simple), datatype classes 0 (integers) and 1 (IEEE floats), any byte order;
  * data layout v1 / v2 / v3, CONTIGUOUS (or compact).  Chunked / filtered datasets raise ``H5Unsupported`` with the
    dataset name -- repack them (``h5repack -l CONTI``) or install ``h5py``; ``sky_embeddings_b200.ingest`` uses
    ``h5py`` automatically when it is importable.

A dataset comes back as a read-only ``numpy.memmap`` over the file: slicing it costs no copy until the bytes are
staged for the host-to-device transfer.

``write_h5`` emits the same byte layout libhdf5 produces for such a file (superblock 0, symbol-table root group,
version-1 object headers, layout v3 contiguous); tests use it to make fixtures, and ``scripts``/users can use it to
export banks without h5py.
"""
from __future__ import annotations

import mmap
import os
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class H5Unsupported(H5Error):
    pass


def _u(buf, off, n):
    return int.from_bytes(buf[off:off + n], "little")


class H5File:
    """Read-only view of an HDF5 file: ``f['cutouts']`` -> numpy memmap, ``f.keys()``, ``len(f['cutouts'])``."""

    def __init__(self, path):
        self.path = path
        self._fh = open(path, "rb")
        size = os.fstat(self._fh.fileno()).st_size
        if size < 64:
            raise H5Error(f"{path}: too small to be an HDF5 file")
        self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        self._size = size
        self._parse_superblock()
        self._links = self._read_group(self._root_addr, self._root_scratch)
        self._cache = {}

    # ------------------------------------------------------------------ superblock
    def _parse_superblock(self):
        mm = self._mm
        off = 0
        while True:     # the signature sits at 0 or at a power of two >= 512 (user block in front)
            if off + 8 <= self._size and mm[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off >= self._size:
                raise H5Error(f"{self.path}: no HDF5 signature found")
        ver = mm[off + 8]
        self._root_scratch = None
        if ver in (0, 1):
            self.O, self.L = mm[off + 13], mm[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = _u(mm, p, self.O)
            p += 4 * self.O                        # base, free-space info, end of file, driver info
            # root group symbol table entry
            self._root_addr = _u(mm, p + self.O, self.O)
            cache_type = _u(mm, p + 2 * self.O, 4)
            if cache_type == 1:
                sp = p + 2 * self.O + 8
                self._root_scratch = (_u(mm, sp, self.O), _u(mm, sp + self.O, self.O))
        elif ver in (2, 3):
            self.O, self.L = mm[off + 9], mm[off + 10]
            p = off + 12
            self.base = _u(mm, p, self.O)
            self._root_addr = _u(mm, p + 3 * self.O, self.O)
        else:
            raise H5Unsupported(f"{self.path}: superblock version {ver}")
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise H5Unsupported(f"{self.path}: offset/length sizes {self.O}/{self.L}")
        if ver >= 2 and self.base == 0:
            self.base = off                        # v2+: addresses are relative to the superblock's own offset
        self._undef_o = (1 << (8 * self.O)) - 1

    def _abs(self, addr):
        if addr == self._undef_o:
            raise H5Error(f"{self.path}: undefined address (dataset never written?)")
        a = addr + self.base
        if a >= self._size:
            raise H5Error(f"{self.path}: address {a:#x} beyond end of file (truncated?)")
        return a

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr):
        """[(type, flags, payload offset, payload size)] of the object header at `addr` (both header versions)."""
        mm = self._mm
        a = self._abs(addr)
        out = []
        if mm[a:a + 4] == b"OHDR":
            if mm[a + 4] != 2:
                raise H5Unsupported(f"{self.path}: object header version {mm[a + 4]}")
            flags = mm[a + 5]
            p = a + 6
            if flags & 0x20:
                p += 16                            # access / modification / change / birth times
            if flags & 0x10:
                p += 4                             # max compact / min dense attributes
            nsz = 1 << (flags & 3)
            chunk = _u(mm, p, nsz)
            p += nsz
            blocks = [(p, p + chunk)]
            track = bool(flags & 4)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, mflags = mm[p], _u(mm, p + 1, 2), mm[p + 3]
                    p += 4 + (2 if track else 0)
                    if mtype == 0x10:
                        ca, cl = self._abs(_u(mm, p, self.O)), _u(mm, p + self.O, self.L)
                        if mm[ca:ca + 4] != b"OCHK":
                            raise H5Error(f"{self.path}: bad continuation block")
                        blocks.append((ca + 4, ca + cl - 4))
                    elif mtype != 0:
                        out.append((mtype, mflags, p, msize))
                    p += msize
            return out
        if mm[a] != 1:
            raise H5Unsupported(f"{self.path}: object header version {mm[a]} at {a:#x}")
        nmsg = _u(mm, a + 2, 2)
        hsize = _u(mm, a + 8, 4)
        blocks = [(a + 16, a + 16 + hsize)]
        while blocks and len(out) < nmsg + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                mtype, msize, mflags = _u(mm, p, 2), _u(mm, p + 2, 2), mm[p + 4]
                p += 8
                if mtype == 0x10:
                    blocks.append((self._abs(_u(mm, p, self.O)), self._abs(_u(mm, p, self.O)) + _u(mm, p + self.O, self.L)))
                elif mtype != 0:
                    out.append((mtype, mflags, p, msize))
                p += msize
        return out

    # ------------------------------------------------------------------ groups
    def _heap_name(self, heap_addr, off):
        mm = self._mm
        h = self._abs(heap_addr)
        if mm[h:h + 4] != b"HEAP":
            raise H5Error(f"{self.path}: bad local heap")
        data = self._abs(_u(mm, h + 8 + 2 * self.L, self.O))
        end = mm.find(b"\0", data + off)
        return mm[data + off:end].decode("utf-8")

    def _walk_btree(self, addr, heap_addr, links):
        mm = self._mm
        a = self._abs(addr)
        if mm[a:a + 4] == b"SNOD":
            n = _u(mm, a + 6, 2)
            p = a + 8
            esz = 2 * self.O + 8 + 16
            for i in range(n):
                e = p + i * esz
                links[self._heap_name(heap_addr, _u(mm, e, self.O))] = _u(mm, e + self.O, self.O)
            return
        if mm[a:a + 4] != b"TREE" or mm[a + 4] != 0:
            raise H5Error(f"{self.path}: bad group B-tree node at {a:#x}")
        used = _u(mm, a + 6, 2)
        p = a + 8 + 2 * self.O + self.L            # past key 0
        for _ in range(used):
            self._walk_btree(_u(mm, p, self.O), heap_addr, links)
            p += self.O + self.L

    def _read_group(self, addr, scratch=None):
        links = {}
        mm = self._mm
        for mtype, _, p, size in self._messages(addr):
            if mtype == 0x11:                      # symbol table: B-tree + local heap
                self._walk_btree(_u(mm, p, self.O), _u(mm, p + self.O, self.O), links)
            elif mtype == 0x06:                    # link message (new-style compact group)
                flags = mm[p + 1]
                q = p + 2
                ltype = 0
                if flags & 0x08:
                    ltype = mm[q]; q += 1
                if flags & 0x04:
                    q += 8
                if flags & 0x10:
                    q += 1
                nsz = 1 << (flags & 3)
                nlen = _u(mm, q, nsz); q += nsz
                name = mm[q:q + nlen].decode("utf-8"); q += nlen
                if ltype == 0:
                    links[name] = _u(mm, q, self.O)
            elif mtype == 0x02:                    # link info: dense storage (fractal heap) is not parsed
                fh = _u(mm, p + 2 + (8 if mm[p + 1] & 1 else 0), self.O)
                if fh != self._undef_o:
                    raise H5Unsupported(f"{self.path}: group with dense link storage (more than 8 links written with "
                                        "libver='latest'); use h5py")
        if not links and scratch is not None:
            self._walk_btree(scratch[0], scratch[1], links)
        return links

    # ------------------------------------------------------------------ datasets
    def keys(self):
        return list(self._links.keys())

    def __contains__(self, name):
        return name in self._links

    def __getitem__(self, name):
        if name not in self._cache:
            if name not in self._links:
                raise KeyError(f"{self.path}: no object named {name!r} (has {sorted(self._links)})")
            self._cache[name] = self._dataset(name, self._links[name])
        return self._cache[name]

    def _dataset(self, name, addr):
        mm = self._mm
        shape = dtype = None
        data_off = None
        nbytes = None
        for mtype, _, p, size in self._messages(addr):
            if mtype == 0x01:                      # dataspace
                ver, rank = mm[p], mm[p + 1]
                q = p + (8 if ver == 1 else 4)
                if ver == 2 and mm[p + 3] == 2:
                    raise H5Unsupported(f"{self.path}:{name}: null dataspace")
                shape = tuple(_u(mm, q + i * self.L, self.L) for i in range(rank))
            elif mtype == 0x03:                    # datatype
                cls, bits0 = mm[p] & 0x0F, mm[p + 1]
                esize = _u(mm, p + 4, 4)
                order = ">" if bits0 & 1 else "<"
                if cls == 1 and esize in (2, 4, 8):
                    dtype = np.dtype(f"{order}f{esize}")
                elif cls == 0 and esize in (1, 2, 4, 8):
                    dtype = np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{esize}")
                else:
                    raise H5Unsupported(f"{self.path}:{name}: datatype class {cls} size {esize}")
            elif mtype == 0x08:                    # data layout
                ver = mm[p]
                if ver in (1, 2):
                    rank, cls = mm[p + 1], mm[p + 2]
                    if cls == 1:
                        data_off = _u(mm, p + 8, self.O)
                    elif cls == 0:
                        q = p + 8 + 4 * rank
                        nbytes = _u(mm, q, 4)
                        data_off = ("compact", q + 4)
                    else:
                        raise H5Unsupported(f"{self.path}:{name}: chunked dataset (layout v{ver}); only contiguous "
                                            "datasets are read without h5py")
                elif ver in (3, 4):
                    cls = mm[p + 1]
                    if cls == 1:
                        data_off, nbytes = _u(mm, p + 2, self.O), _u(mm, p + 2 + self.O, self.L)
                    elif cls == 0:
                        nbytes = _u(mm, p + 2, 2)
                        data_off = ("compact", p + 4)
                    else:
                        raise H5Unsupported(f"{self.path}:{name}: chunked / virtual dataset (layout class {cls}); only "
                                            "contiguous datasets are read without h5py")
                else:
                    raise H5Unsupported(f"{self.path}:{name}: data layout version {ver}")
            elif mtype == 0x0B:
                raise H5Unsupported(f"{self.path}:{name}: filtered (compressed) dataset; only plain contiguous data is "
                                    "read without h5py")
        if shape is None or dtype is None or data_off is None:
            raise H5Error(f"{self.path}:{name}: not a dataset (dataspace / datatype / layout message missing)")
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if isinstance(data_off, tuple):
            return np.frombuffer(mm, dtype=dtype, count=count, offset=data_off[1]).reshape(shape)
        if count == 0:
            return np.zeros(shape, dtype=dtype)
        a = self._abs(data_off)
        if a + count * dtype.itemsize > self._size:
            raise H5Error(f"{self.path}:{name}: data runs past the end of the file (truncated?)")
        return np.memmap(self.path, dtype=dtype, mode="r", offset=a, shape=shape)

    def close(self):
        self._cache.clear()
        try:
            self._mm.close()
        except BufferError:      # a compact dataset still references the map
            pass
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---------------------------------------------------------------------------------------------------------------
# writer: the byte layout libhdf5 emits for create_dataset(name, shape, dtype) with default (earliest) settings
# ---------------------------------------------------------------------------------------------------------------
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, payload, flags=0):
    payload = _pad8(payload)
    return struct.pack("<HHB3x", mtype, len(payload), flags) + payload


def _dtype_msg(dt):
    dt = np.dtype(dt)
    be = 1 if dt.byteorder == ">" else 0
    if dt.kind == "f":
        exp, man = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[dt.itemsize]
        bits = bytes([be | 0x20, 8 * dt.itemsize - 1, 0])
        prop = struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, man, exp, 0, man, (1 << (exp - 1)) - 1)
        return bytes([0x11]) + bits + struct.pack("<I", dt.itemsize) + prop
    if dt.kind in "iu":
        bits = bytes([be | (0x08 if dt.kind == "i" else 0), 0, 0])
        return bytes([0x10]) + bits + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise H5Unsupported(f"cannot write dtype {dt}")


def write_h5(path, datasets, userblock=0):
    """Write ``{name: array}`` as contiguous datasets of one root group (at most 8 names: one symbol node).
    ``userblock``: bytes reserved in front of the superblock (0 or a power of two >= 512), as MATLAB / some tools do."""
    names = sorted(datasets)                       # symbol nodes are ordered by name
    if not 1 <= len(names) <= 8:
        raise H5Unsupported("write_h5 writes 1..8 datasets in the root group")
    if userblock not in (0,) and (userblock < 512 or userblock & (userblock - 1)):
        raise H5Error("userblock must be 0 or a power of two >= 512")
    arrays = {n: np.ascontiguousarray(datasets[n]) for n in names}
    O = L = 8
    # layout of the metadata (addresses relative to the base address = start of the superblock)
    sb_size = 8 + 8 + 4 + 4 + 4 * O + (2 * O + 8 + 16)          # 96
    root_oh = sb_size
    root_oh_size = 16 + 8 + 16 + 8                              # prefix + symbol-table message + a NIL message
    heap_hdr = root_oh + root_oh_size
    heap_data_size = 8 + sum(len(_pad8(n.encode() + b"\0")) for n in names)
    heap_data_size += -heap_data_size % 8 + 16                  # room for the free-list block
    heap_data = heap_hdr + 8 + 2 * L + O
    btree = heap_data + heap_data_size
    btree_size = 8 + 2 * O + (2 * 16 + 1) * L + 2 * 16 * O      # node sized for internal K = 16
    snod = btree + btree_size
    snod_size = 8 + 8 * (2 * O + 8 + 16)                        # leaf K = 4 -> 8 entries
    p = snod + snod_size
    # heap contents
    heap = bytearray(b"\0" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode() + b"\0")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, heap_data_size - free_off)    # free block: next = 1 (none), size
    heap = bytes(heap).ljust(heap_data_size, b"\0")
    # dataset object headers, then raw data (8-byte aligned)
    oh_addr, oh_bytes = {}, {}
    for n in names:
        a = arrays[n]
        msgs = [_msg(0x01, struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", d) for d in a.shape)),
                _msg(0x03, _dtype_msg(a.dtype), flags=1),
                _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0)),                       # fill value v2: alloc late, never written
                None,                                                               # layout, needs the data address
                _msg(0x12, struct.pack("<B3xI", 1, 0))]                             # modification time v1
        oh_addr[n] = p
        oh_bytes[n] = msgs
        p += 16 + sum(len(m) for m in msgs if m) + 8 + 2 + O + L + (-(2 + O + L) % 8)
    data_addr = {}
    for n in names:
        p += -p % 8
        data_addr[n] = p
        p += arrays[n].nbytes
    eof = p
    with open(path, "wb") as f:
        f.write(b"\0" * userblock)
        sb = SIGNATURE + bytes([0, 0, 0, 0, 0, O, L, 0]) + struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", userblock, UNDEF, userblock + eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_oh, 1, 0) + struct.pack("<QQ", btree, heap_hdr)
        assert len(sb) == sb_size
        f.write(sb)
        f.write(struct.pack("<BBHII4x", 1, 0, 2, 1, root_oh_size - 16))
        f.write(_msg(0x11, struct.pack("<QQ", btree, heap_hdr)))
        f.write(_msg(0x00, b""))
        f.write(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", heap_data_size, free_off, heap_data))
        f.write(heap)
        node = b"TREE" + bytes([0, 0]) + struct.pack("<H", 1) + struct.pack("<QQ", UNDEF, UNDEF)
        node += struct.pack("<QQQ", 0, snod, name_off[names[-1]])
        f.write(node.ljust(btree_size, b"\0"))
        s = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(names))
        for n in names:
            s += struct.pack("<QQII16x", name_off[n], oh_addr[n], 0, 0)
        f.write(s.ljust(snod_size, b"\0"))
        for n in names:
            msgs = list(oh_bytes[n])
            msgs[3] = _msg(0x08, struct.pack("<BBQQ", 3, 1, data_addr[n], arrays[n].nbytes))
            body = b"".join(msgs)
            assert f.tell() == userblock + oh_addr[n]
            f.write(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)
        for n in names:
            f.write(b"\0" * (userblock + data_addr[n] - f.tell()))
            f.write(arrays[n].tobytes())
    return path

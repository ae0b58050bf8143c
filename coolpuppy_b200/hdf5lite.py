"""Minimal read-only HDF5 decoder, enough for ``.cool`` files.

No h5py / libhdf5 exists in this image, and the pile-up path only needs a
handful of 1-D chunked datasets out of a cooler, so this module decodes exactly
the subset of the HDF5 file format that cooler writes:

* superblock version 0 (8-byte offsets and lengths),
* version-1 object headers (with continuation blocks),
* "old style" groups: symbol-table message -> v1 B-tree -> ``SNOD`` nodes
  + local heap for the link names,
* datasets with dataspace v1/v2, datatype classes 0 (integer), 1 (float),
  3 (fixed string) and 8 (enum over an integer base), data layout v3
  contiguous or chunked (v1 chunk B-tree), filter pipeline v1/v2 with
  shuffle (id 2) and deflate (id 1),
* version-1 attribute messages with scalar numeric / fixed-string values
  (variable-length strings are reported as ``None``).

Everything is little-endian, which is what h5py writes on every platform
cooler runs on.  The layout facts were established in SURVEY.md Appendix B.
"""
from __future__ import annotations

import struct
import os
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5FormatError(ValueError):
    pass


class _Dataset:
    def __init__(self, f, name, msgs):
        self._f = f
        self.name = name
        self.shape = None
        self.dtype = None
        self.enum = None
        self._layout = None
        self._filters = []
        self.attrs = {}
        for mtype, body in msgs:
            if mtype == 0x01:
                self.shape = _parse_dataspace(body)
            elif mtype == 0x03:
                self.dtype, self.enum, _ = _parse_datatype(body, 0)
            elif mtype == 0x08:
                self._layout = _parse_layout(body)
            elif mtype == 0x0B:
                self._filters = _parse_filters(body)
            elif mtype == 0x0C:
                k, v = _parse_attribute(body)
                self.attrs[k] = v
        if self.shape is None or self.dtype is None or self._layout is None:
            raise HDF5FormatError(f"{name}: not a dataset")

    def __len__(self):
        return self.shape[0] if self.shape else 1

    def read(self):
        """Whole dataset as a numpy array (1-D datasets only)."""
        if len(self.shape) != 1:
            raise HDF5FormatError(f"{self.name}: only 1-D datasets are supported")
        n = self.shape[0]
        out = np.empty(n, dtype=self.dtype)
        kind = self._layout[0]
        buf = self._f._buf
        if kind == "contiguous":
            _, addr, size = self._layout
            if addr == _UNDEF or n == 0:
                out[:] = 0
            else:
                out[:] = np.frombuffer(buf, dtype=self.dtype, count=n, offset=addr)
        elif kind == "compact":
            _, raw = self._layout
            out[:] = np.frombuffer(raw, dtype=self.dtype, count=n)
        else:
            _, btree, chunk_dims = self._layout
            chunk = chunk_dims[0]
            itemsize = self.dtype.itemsize
            out_view = out.view(np.uint8)
            out[:] = 0
            if btree != _UNDEF:
                filters = list(reversed(list(enumerate(self._filters))))

                def decode(entry):
                    nbytes, mask, off, addr = entry
                    raw = bytes(buf[addr : addr + nbytes])
                    for i, (fid, cd) in filters:
                        if mask & (1 << i):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            es = cd[0] if cd else itemsize
                            a = np.frombuffer(raw, dtype=np.uint8)
                            ne = a.size // es
                            body = a[: ne * es].reshape(es, ne).T.reshape(-1)
                            raw = body.tobytes() + a[ne * es :].tobytes()
                        else:
                            raise HDF5FormatError(f"{self.name}: unsupported filter id {fid}")
                    start = off[0]
                    count = min(chunk, n - start)
                    if count > 0:  # chunks cover disjoint ranges: safe to fill from several threads
                        out_view[start * itemsize : (start + count) * itemsize] = np.frombuffer(
                            raw, dtype=np.uint8, count=count * itemsize
                        )

                entries = list(self._f._iter_chunks(btree, 2))
                workers = min(len(entries) // 4, os.cpu_count() or 1, 16)
                if workers >= 2:  # zlib and the numpy copies release the GIL
                    from concurrent.futures import ThreadPoolExecutor

                    step = max(1, len(entries) // (workers * 4))  # a few batches per thread: chunks can be tiny
                    batches = [entries[i : i + step] for i in range(0, len(entries), step)]
                    with ThreadPoolExecutor(max_workers=workers) as pool:
                        list(pool.map(lambda batch: [decode(e) for e in batch], batches))
                else:
                    for e in entries:
                        decode(e)
        return out

    def __getitem__(self, key):
        return self.read()[key]


class _Group:
    def __init__(self, f, name, btree, heap, msgs):
        self._f = f
        self.name = name
        self._links = f._read_links(btree, heap)
        self.attrs = {}
        for mtype, body in msgs:
            if mtype == 0x0C:
                k, v = _parse_attribute(body)
                self.attrs[k] = v

    def keys(self):
        return list(self._links.keys())

    def __contains__(self, key):
        return key in self._links

    def __getitem__(self, key):
        node = self
        for part in key.strip("/").split("/"):
            if part not in node._links:
                raise KeyError(f"{key!r} not in {self.name!r}")
            node = self._f._open(node._links[part], f"{node.name.rstrip('/')}/{part}")
        return node


class File(_Group):
    """``File(path)["bins/start"].read()`` -> numpy array."""

    def __init__(self, path):
        self.filename = str(path)
        with open(path, "rb") as fh:
            self._buf = memoryview(fh.read())
        b = self._buf
        if bytes(b[:8]) != _SIG:
            raise HDF5FormatError(f"{path}: not an HDF5 file")
        if b[8] != 0:
            raise HDF5FormatError(f"{path}: superblock version {b[8]} not supported (need 0)")
        if b[13] != 8 or b[14] != 8:
            raise HDF5FormatError(f"{path}: only 8-byte offsets/lengths are supported")
        # superblock v0: root symbol-table entry starts at byte 56; its
        # object-header address is the second u64 of the entry.
        root_addr = struct.unpack_from("<Q", b, 64)[0]
        self._cache = {}
        msgs = self._read_object_header(root_addr)
        st = [body for t, body in msgs if t == 0x11]
        if not st:
            raise HDF5FormatError("root object has no symbol table")
        btree, heap = struct.unpack_from("<QQ", st[0], 0)
        _Group.__init__(self, self, "/", btree, heap, msgs)

    # -- object headers -------------------------------------------------------
    def _read_object_header(self, addr):
        b = self._buf
        ver, _, nmsgs, _refcnt, hdr_size = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise HDF5FormatError(f"object header version {ver} at {addr} not supported")
        blocks = [(addr + 16, hdr_size)]
        msgs = []
        while blocks and len(msgs) < nmsgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(msgs) < nmsgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = bytes(b[pos + 8 : pos + 8 + msize])
                pos += 8 + msize
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr, clen))
                msgs.append((mtype, body))
        return msgs

    def _open(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._read_object_header(addr)
        st = [body for t, body in msgs if t == 0x11]
        if st:
            btree, heap = struct.unpack_from("<QQ", st[0], 0)
            obj = _Group(self, name, btree, heap, msgs)
        else:
            obj = _Dataset(self, name, msgs)
        self._cache[addr] = obj
        return obj

    # -- groups ---------------------------------------------------------------
    def _read_links(self, btree, heap):
        b = self._buf
        if bytes(b[heap : heap + 4]) != b"HEAP":
            raise HDF5FormatError("bad local heap signature")
        heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
        links = {}

        def walk(node):
            sig = bytes(b[node : node + 4])
            if sig == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
                if ntype != 0:
                    raise HDF5FormatError("expected a group B-tree")
                pos = node + 24
                for i in range(used):
                    child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0]
                    walk(child)
            elif sig == b"SNOD":
                nsyms = struct.unpack_from("<H", b, node + 6)[0]
                for i in range(nsyms):
                    name_off, ohdr = struct.unpack_from("<QQ", b, node + 8 + i * 40)
                    p = heap_data + name_off
                    q = p
                    while b[q] != 0:
                        q += 1
                    links[bytes(b[p:q]).decode()] = ohdr
            else:
                raise HDF5FormatError(f"unexpected node signature {sig!r}")

        walk(btree)
        return links

    # -- chunk index ----------------------------------------------------------
    def _iter_chunks(self, node, ndims):
        b = self._buf
        if bytes(b[node : node + 4]) != b"TREE":
            raise HDF5FormatError("bad chunk B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
        if ntype != 1:
            raise HDF5FormatError("expected a chunk B-tree")
        keysize = 8 + 8 * ndims
        pos = node + 24
        for i in range(used):
            kpos = pos + i * (keysize + 8)
            nbytes, mask = struct.unpack_from("<II", b, kpos)
            off = struct.unpack_from(f"<{ndims}Q", b, kpos + 8)
            child = struct.unpack_from("<Q", b, kpos + keysize)[0]
            if level == 0:
                yield nbytes, mask, off, child
            else:
                yield from self._iter_chunks(child, ndims)


# -- message parsers -----------------------------------------------------------
def _parse_dataspace(body):
    ver, rank, flags = struct.unpack_from("<BBB", body, 0)
    if ver == 1:
        pos = 8
    elif ver == 2:
        pos = 4
    else:
        raise HDF5FormatError(f"dataspace version {ver} not supported")
    return tuple(struct.unpack_from(f"<{rank}Q", body, pos)) if rank else ()


def _parse_datatype(body, pos):
    """Returns (numpy dtype, enum mapping or None, bytes consumed)."""
    cls_ver, b0, b1, b2, size = struct.unpack_from("<BBBBI", body, pos)
    cls = cls_ver & 0x0F
    if cls == 0:
        if b0 & 1:
            raise HDF5FormatError("big-endian integers not supported")
        signed = bool(b0 & 0x08)
        return np.dtype(f"<{'i' if signed else 'u'}{size}"), None, 8 + 4
    if cls == 1:
        if b0 & 1:
            raise HDF5FormatError("big-endian floats not supported")
        return np.dtype(f"<f{size}"), None, 8 + 12
    if cls == 3:
        return np.dtype(f"S{size}"), None, 8
    if cls == 8:
        nmemb = b0 | (b1 << 8)
        base, _, used = _parse_datatype(body, pos + 8)
        p = pos + 8 + used
        names = []
        for _ in range(nmemb):
            q = body.index(b"\x00", p)
            names.append(body[p:q].decode())
            ln = q - p + 1
            p += (ln + 7) // 8 * 8
        vals = np.frombuffer(body, dtype=base, count=nmemb, offset=p)
        return base, dict(zip((int(v) for v in vals), names)), p + nmemb * base.itemsize - pos
    if cls == 9:
        # variable length (strings in attributes): value lives in the global heap
        return None, None, 8
    raise HDF5FormatError(f"datatype class {cls} not supported")


def _parse_layout(body):
    ver, cls = struct.unpack_from("<BB", body, 0)
    if ver != 3:
        raise HDF5FormatError(f"data layout version {ver} not supported")
    if cls == 0:
        size = struct.unpack_from("<H", body, 2)[0]
        return ("compact", body[4 : 4 + size])
    if cls == 1:
        addr, size = struct.unpack_from("<QQ", body, 2)
        return ("contiguous", addr, size)
    if cls == 2:
        ndims = body[2]
        btree = struct.unpack_from("<Q", body, 3)[0]
        dims = struct.unpack_from(f"<{ndims}I", body, 11)
        if ndims != 2:
            raise HDF5FormatError("only 1-D chunked datasets are supported")
        return ("chunked", btree, dims)
    raise HDF5FormatError(f"layout class {cls} not supported")


def _parse_filters(body):
    ver, nfilters = struct.unpack_from("<BB", body, 0)
    out = []
    if ver == 1:
        pos = 8
        for _ in range(nfilters):
            fid, name_len, _flags, ncd = struct.unpack_from("<HHHH", body, pos)
            pos += 8 + (name_len + 7) // 8 * 8
            cd = struct.unpack_from(f"<{ncd}I", body, pos)
            pos += 4 * ncd + (4 if ncd % 2 else 0)
            out.append((fid, cd))
    elif ver == 2:
        pos = 2
        for _ in range(nfilters):
            fid = struct.unpack_from("<H", body, pos)[0]
            pos += 2
            name_len = 0
            if fid >= 256:
                name_len = struct.unpack_from("<H", body, pos)[0]
                pos += 2
            _flags, ncd = struct.unpack_from("<HH", body, pos)
            pos += 4 + name_len
            cd = struct.unpack_from(f"<{ncd}I", body, pos)
            pos += 4 * ncd
            out.append((fid, cd))
    else:
        raise HDF5FormatError(f"filter pipeline version {ver} not supported")
    return out


def _parse_attribute(body):
    ver = body[0]
    if ver != 1:
        return f"<attr v{ver}>", None
    name_size, dt_size, ds_size = struct.unpack_from("<HHH", body, 2)
    pos = 8
    name = body[pos : pos + name_size].split(b"\x00")[0].decode()
    pos += (name_size + 7) // 8 * 8
    try:
        dtype, _, _ = _parse_datatype(body, pos)
    except HDF5FormatError:
        dtype = None
    pos += (dt_size + 7) // 8 * 8
    try:
        shape = _parse_dataspace(body[pos : pos + ds_size])
    except HDF5FormatError:
        shape = None
    pos += (ds_size + 7) // 8 * 8
    if dtype is None or shape is None:
        return name, None
    n = int(np.prod(shape)) if shape else 1
    if len(body) < pos + n * dtype.itemsize:
        return name, None
    val = np.frombuffer(body, dtype=dtype, count=n, offset=pos)
    if dtype.kind == "S":
        val = [v.split(b"\x00")[0].decode() for v in val]
    if not shape:
        val = val[0]
        if isinstance(val, np.generic):
            val = val.item()
    return name, val

"""Cooler-compatible matrix access without ``cooler``/``h5py``.

The reference reads its input through ``cooler.Cooler`` (``coolpup.py:838,
922-925, 1053-1055, 1082-1098``).  Neither ``cooler`` nor an HDF5 library is
present in this image, so this module provides the small slice of that API the
pile-up path touches, backed either by a ``.cool`` file decoded with
:mod:`hdf5lite` (:class:`Cooler`) or by in-memory arrays (:class:`MemCooler`,
used for the synthetic benchmark genomes).

Only single-resolution, symmetric-upper coolers with a fixed bin size are
supported -- that is the only storage mode coolpuppy accepts anyway.

API mirrored (cooler 0.9 semantics, restated from the cooler documentation):

* ``binsize``, ``chromnames``, ``chromsizes``, ``filename``, ``info``
* ``offset(chrom_or_region)``, ``extent(region)``
* ``bins()`` -> selector with ``.columns``, ``[col]`` / ``[:]`` and ``.fetch``
* ``matrix(sparse=True, balance=<name|False|True>).fetch(region1, region2)``
  -> ``scipy.sparse.coo_matrix`` with the lower triangle mirrored in and
  ``value = (w[row] * w[col]) * count`` when balanced.
* ``region_csr(lo, hi)`` -- *not* in cooler: the symmetric-filled raw-count CSR
  of a square bin range as ``(indptr int32, indices int32, counts int32)``,
  which is what the CUDA path consumes.
"""
from __future__ import annotations

import os
import re

import numpy as np
import pandas as pd

from . import hdf5lite

__all__ = ["Cooler", "MemCooler", "ChromCooler", "parse_region", "is_cooler"]


def parse_region(region, chromsizes):
    """(chrom, start, end) from a UCSC string, a 1-/3-tuple or a frame row."""
    if isinstance(region, str):
        m = re.fullmatch(r"([^:]+)(?::([\d,]*)-([\d,]*))?", region)
        if m is None:
            raise ValueError(f"cannot parse region {region!r}")
        chrom = m.group(1)
        start = int(m.group(2).replace(",", "")) if m.group(2) else 0
        end = int(m.group(3).replace(",", "")) if m.group(3) else int(chromsizes[chrom])
    else:
        if isinstance(region, pd.Series):
            vals = list(region.iloc[:3]) if len(region) >= 3 else list(region)
        else:
            vals = list(region)
        if len(vals) == 1:
            vals = [vals[0], None, None]
        chrom, start, end = vals[0], vals[1], vals[2]
        start = 0 if start is None else int(start)
        end = int(chromsizes[chrom]) if end is None else int(end)
    if chrom not in chromsizes.index:
        raise ValueError(f"unknown chromosome {chrom!r}")
    if start < 0 or end > int(chromsizes[chrom]) or start > end:
        raise ValueError(f"region {chrom}:{start}-{end} out of bounds")
    return chrom, start, end


class _ColumnSelector:
    def __init__(self, clr, col):
        self._clr = clr
        self._col = col

    def fetch(self, region):
        lo, hi = self._clr.extent(region)
        return pd.Series(self._clr._bin_column(self._col)[lo:hi], index=np.arange(lo, hi), name=self._col)

    def __getitem__(self, key):
        return pd.Series(self._clr._bin_column(self._col)[key], name=self._col)

    @property
    def values(self):
        return self._clr._bin_column(self._col)


class _BinSelector:
    def __init__(self, clr):
        self._clr = clr

    @property
    def columns(self):
        return pd.Index(self._clr._bin_columns())

    def __getitem__(self, key):
        if isinstance(key, str):
            if key not in self._clr._bin_columns():
                raise KeyError(key)
            return _ColumnSelector(self._clr, key)
        if isinstance(key, (list, tuple)) and all(isinstance(k, str) for k in key):
            return pd.DataFrame({k: self._clr._bin_column(k) for k in key})
        frame = pd.DataFrame({k: self._clr._bin_column(k) for k in self._clr._bin_columns()})
        return frame[key]

    def fetch(self, region):
        lo, hi = self._clr.extent(region)
        return self[:].iloc[lo:hi]


class _MatrixSelector:
    def __init__(self, clr, balance, sparse):
        self._clr = clr
        self._balance = balance
        self._sparse = sparse

    def fetch(self, region1, region2=None):
        from scipy import sparse as sp

        clr = self._clr
        if region2 is None:
            region2 = region1
        i0, i1 = clr.extent(region1)
        j0, j1 = clr.extent(region2)
        row, col, val = clr._query_rect(i0, i1, j0, j1)
        mat = sp.coo_matrix((val, (row - i0, col - j0)), shape=(i1 - i0, j1 - j0))
        if self._balance:
            name = "weight" if self._balance is True else self._balance
            w = clr._bin_column(name).astype(np.float64)
            if clr._divisive(name):
                w = 1.0 / w
            mat.data = w[i0:i1][mat.row] * w[j0:j1][mat.col] * mat.data
        if self._sparse:
            return mat
        return mat.toarray()


class _CoolerBase:
    """Shared query logic; subclasses provide the arrays."""

    filename = None

    # -- to be provided -------------------------------------------------------
    binsize: int
    chromnames: list
    chromsizes: pd.Series
    _chrom_offset: np.ndarray  # int64 [nchroms+1]
    _bin1_offset: np.ndarray  # int64 [nbins+1]
    _bin1: np.ndarray
    _bin2: np.ndarray
    _count: np.ndarray

    def _bin_columns(self):
        raise NotImplementedError

    def _bin_column(self, name):
        raise NotImplementedError

    def _divisive(self, name):
        return False

    # -- cooler API -----------------------------------------------------------
    @property
    def info(self):
        return {"bin-size": self.binsize, "nbins": int(self._chrom_offset[-1]), "nchroms": len(self.chromnames)}

    def offset(self, region):
        if isinstance(region, str) and region in self.chromsizes.index:
            return int(self._chrom_offset[self.chromnames.index(region)])
        return self.extent(region)[0]

    def extent(self, region):
        chrom, start, end = parse_region(region, self.chromsizes)
        off = int(self._chrom_offset[self.chromnames.index(chrom)])
        lo = off + start // self.binsize
        hi = off + -(-end // self.binsize)
        return lo, hi

    def bins(self):
        return _BinSelector(self)

    def matrix(self, field=None, balance=True, sparse=False, **_ignored):
        if field not in (None, "count"):
            raise NotImplementedError("only the 'count' field is supported")
        return _MatrixSelector(self, balance, sparse)

    # -- pixel queries --------------------------------------------------------
    def _upper_pixels(self, lo, hi):
        """Stored (upper-triangle) pixels with both bins in [lo, hi)."""
        p0, p1 = int(self._bin1_offset[lo]), int(self._bin1_offset[hi])
        b1 = self._bin1[p0:p1]
        b2 = self._bin2[p0:p1]
        keep = b2 < hi  # bin2 >= bin1 >= lo holds for symmetric-upper storage
        return b1[keep], b2[keep], self._count[p0:p1][keep]

    def _query_rect(self, i0, i1, j0, j1):
        lo, hi = min(i0, j0), max(i1, j1)
        b1, b2, c = self._upper_pixels(lo, hi)
        off = b1 != b2
        row = np.concatenate([b1, b2[off]])
        col = np.concatenate([b2, b1[off]])
        val = np.concatenate([c, c[off]])
        keep = (row >= i0) & (row < i1) & (col >= j0) & (col < j1)
        return row[keep].astype(np.int64), col[keep].astype(np.int64), val[keep]

    def region_upper_csr(self, lo, hi):
        """Upper-triangle CSR rows of the bin range [lo, hi) exactly as the cooler stores them.

        Returns ``(indptr int32[nb+1], cols int32[n], counts int32[n])`` with region-relative columns; pixels that
        leave the region (column >= nb: trans contacts, or cis beyond a view arm) are kept and have to be dropped
        by the consumer (``pup_region_create_upper`` does).  No sorting or filtering happens on the host.
        """
        p0, p1 = int(self._bin1_offset[lo]), int(self._bin1_offset[hi])
        if p1 - p0 >= 2**30:
            raise ValueError("region has more than 2^30 stored upper-triangle pixels")
        indptr = (self._bin1_offset[lo : hi + 1] - p0).astype(np.int32)
        cols = np.minimum(self._bin2[p0:p1] - lo, 2**31 - 1).astype(np.int32)
        cnt = self._count[p0:p1]
        if not np.issubdtype(cnt.dtype, np.integer):
            raise NotImplementedError("floating-point pixel counts are not supported by the CUDA path")
        if cols.size > 1:
            # the format requires pixels sorted by (bin1, bin2); some files in the wild (e.g. the reference's
            # CN.mm9.1000kb.cool test fixture) have out-of-order tails inside a row: sort those rows here
            desc = np.diff(cols) <= 0
            if indptr.size > 2:
                inner = indptr[1:-1]
                desc[inner[(inner > 0) & (inner < cols.size)] - 1] = False  # row boundaries may descend
            if desc.any():
                rows = np.repeat(np.arange(hi - lo, dtype=np.int64), np.diff(indptr))
                order = np.lexsort((cols, rows))
                cols, cnt = cols[order], cnt[order]
        return indptr, cols, np.ascontiguousarray(cnt, dtype=np.int32)

    def region_csr(self, lo, hi):
        """Symmetric-filled raw-count CSR of the square bin range [lo, hi).

        Returns ``(indptr int32[nb+1], indices int32[nnz], counts int32[nnz])``
        with column indices sorted within each row (region-relative bins).
        """
        b1, b2, c = self._upper_pixels(lo, hi)
        nb = hi - lo
        off = b1 != b2
        row = np.concatenate([b1, b2[off]]) - lo
        col = np.concatenate([b2, b1[off]]) - lo
        val = np.concatenate([c, c[off]])
        if row.size >= 2**31:
            raise ValueError("region has more than 2^31 stored pixels")
        order = np.lexsort((col, row))
        counts = np.bincount(row, minlength=nb)
        indptr = np.zeros(nb + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])
        cnt = val[order]
        if not np.issubdtype(cnt.dtype, np.integer):
            raise NotImplementedError("floating-point pixel counts are not supported by the CUDA path")
        return indptr.astype(np.int32), col[order].astype(np.int32), cnt.astype(np.int32)


class Cooler(_CoolerBase):
    """Read-only view of a single-resolution ``.cool`` file."""

    def __init__(self, path):
        path = str(path)
        if "::" in path:
            path, group = path.split("::", 1)
        else:
            group = "/"
        self.filename = path
        self._h5 = hdf5lite.File(path)
        self._root = self._h5 if group.strip("/") == "" else self._h5[group]
        attrs = self._root.attrs
        if attrs.get("storage-mode") not in (None, "symmetric-upper"):
            raise NotImplementedError("only symmetric-upper coolers are supported")
        self.binsize = int(attrs["bin-size"])
        names = self._root["chroms/name"].read()
        self.chromnames = [n.split(b"\x00")[0].decode() for n in names]
        lengths = self._root["chroms/length"].read().astype(np.int64)
        self.chromsizes = pd.Series(lengths, index=pd.Index(self.chromnames, name="name"), name="length")
        self._chrom_offset = self._root["indexes/chrom_offset"].read().astype(np.int64)
        self._bin1_offset = self._root["indexes/bin1_offset"].read().astype(np.int64)
        self._cols = {}
        self._pix = None

    # pixels are decoded lazily, once
    def _load_pixels(self):
        if self._pix is None:
            self._pix = (
                self._root["pixels/bin1_id"].read(),
                self._root["pixels/bin2_id"].read(),
                self._root["pixels/count"].read(),
            )
        return self._pix

    @property
    def _bin1(self):
        return self._load_pixels()[0]

    @property
    def _bin2(self):
        return self._load_pixels()[1]

    @property
    def _count(self):
        return self._load_pixels()[2]

    def _bin_columns(self):
        stored = list(self._root["bins"].keys())
        return stored + [k for k in self._cols if k not in stored]

    def add_bin_column(self, name, values):
        """Attach an in-memory bin column (e.g. computed coverage); the file is not modified."""
        values = np.asarray(values)
        if values.shape[0] != int(self._chrom_offset[-1]):
            raise ValueError(f"bin column {name!r} has {values.shape[0]} rows, expected {int(self._chrom_offset[-1])}")
        self._cols[name] = values

    def _bin_column(self, name):
        if name not in self._cols:
            ds = self._root["bins"][name]
            arr = ds.read()
            if name == "chrom" and ds.enum is not None:
                arr = pd.Categorical.from_codes(arr, categories=[ds.enum[i] for i in sorted(ds.enum)])
                arr = np.asarray(arr.astype(str))
            self._cols[name] = arr
        return self._cols[name]

    def _divisive(self, name):
        return bool(self._root["bins"][name].attrs.get("divisive_weights", False))


class MemCooler(_CoolerBase):
    """In-memory cooler built from arrays (synthetic genomes, tests).

    ``bin1``/``bin2`` are global bin ids of the stored upper-triangle pixels,
    sorted by ``(bin1, bin2)``; ``bin_columns`` maps extra bin-table column
    names (``weight``, ``cov_tot_raw`` ...) to arrays of length ``nbins``.
    """

    def __init__(self, chromsizes, binsize, bin1, bin2, count, bin_columns=None, filename="<memory>.cool"):
        self.filename = filename
        self.binsize = int(binsize)
        if isinstance(chromsizes, dict):
            chromsizes = pd.Series(chromsizes)
        self.chromnames = [str(c) for c in chromsizes.index]
        self.chromsizes = pd.Series(
            np.asarray(chromsizes.values, dtype=np.int64), index=pd.Index(self.chromnames, name="name"), name="length"
        )
        nb = -(-self.chromsizes.values // self.binsize)
        self._chrom_offset = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
        nbins = int(self._chrom_offset[-1])
        bin1 = np.asarray(bin1)
        bin2 = np.asarray(bin2)
        if bin1.size and (np.any(bin1 > bin2) or np.any(np.diff(bin1) < 0)):
            raise ValueError("pixels must be upper-triangular and sorted by bin1")
        self._bin1 = bin1
        self._bin2 = bin2
        self._count = np.asarray(count)
        self._bin1_offset = np.searchsorted(bin1, np.arange(nbins + 1), side="left").astype(np.int64)
        starts = np.concatenate([np.arange(n, dtype=np.int64) * self.binsize for n in nb])
        chrom_col = np.repeat(np.asarray(self.chromnames, dtype=object), nb)
        ends = np.minimum(starts + self.binsize, np.repeat(self.chromsizes.values, nb))
        self._cols = {"chrom": chrom_col, "start": starts, "end": ends}
        for k, v in (bin_columns or {}).items():
            v = np.asarray(v)
            if v.shape[0] != nbins:
                raise ValueError(f"bin column {k!r} has {v.shape[0]} rows, expected {nbins}")
            self._cols[k] = v

    def _bin_columns(self):
        return list(self._cols)

    def _bin_column(self, name):
        return self._cols[name]

    def add_bin_column(self, name, values):
        self._cols[name] = np.asarray(values)


class ChromCooler(_CoolerBase):
    """In-memory cooler that keeps every chromosome's cis pixels as the region-relative upper-triangle CSR the CUDA
    path consumes (``int32 indptr[nb+1], col[nnz], count[nnz]``), so that ``region_upper_csr`` of a whole-chromosome
    view is a zero-copy lookup -- with the arrays in pinned memory (``pin=True``, or ``pin="lazy"``: a chromosome is
    pinned the first time it is read) the host->device copies of the region pipeline are truly asynchronous.  Trans pixels are not represented.  Sub-chromosome views and the cooler-style
    ``matrix().fetch`` work too (they build global pixel ids on demand).
    """

    def __init__(self, chromsizes, binsize, regions, bin_columns=None, filename="<memory>.cool", pin=False):
        self.filename = filename
        self._lazy_pin = pin == "lazy"
        self._pinned = set()
        pin = pin is True
        self.binsize = int(binsize)
        if isinstance(chromsizes, dict):
            chromsizes = pd.Series(chromsizes)
        self.chromnames = [str(c) for c in chromsizes.index]
        self.chromsizes = pd.Series(
            np.asarray(chromsizes.values, dtype=np.int64), index=pd.Index(self.chromnames, name="name"), name="length"
        )
        nb = -(-self.chromsizes.values // self.binsize)
        self._chrom_offset = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
        nbins = int(self._chrom_offset[-1])
        self._keep = []  # pinned torch tensors backing the numpy views

        def hold(a, dtype):
            a = np.ascontiguousarray(a, dtype=dtype)
            if not pin:
                return a
            import torch

            t = torch.from_numpy(a).pin_memory()
            self._keep.append(t)
            return t.numpy()

        self._regions = {}
        offs = np.zeros(nbins + 1, dtype=np.int64)
        total = 0
        for ci, c in enumerate(self.chromnames):
            n = int(nb[ci])
            if c in regions:
                ip, col, cnt = regions[c]
                ip = np.asarray(ip)
                if ip.shape[0] != n + 1:
                    raise ValueError(f"{c}: indptr has {ip.shape[0]} entries, expected {n + 1}")
                self._regions[c] = (hold(ip, np.int32), hold(col, np.int32), hold(cnt, np.int32))
            else:
                self._regions[c] = (np.zeros(n + 1, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
            lo = int(self._chrom_offset[ci])
            offs[lo : lo + n + 1] = total + self._regions[c][0].astype(np.int64)
            total += int(self._regions[c][0][-1])
        self._bin1_offset = offs
        starts = np.concatenate([np.arange(n, dtype=np.int64) * self.binsize for n in nb])
        chrom_col = np.repeat(np.asarray(self.chromnames, dtype=object), nb)
        ends = np.minimum(starts + self.binsize, np.repeat(self.chromsizes.values, nb))
        self._cols = {"chrom": chrom_col, "start": starts, "end": ends}
        for k, v in (bin_columns or {}).items():
            v = np.asarray(v)
            if v.shape[0] != nbins:
                raise ValueError(f"bin column {k!r} has {v.shape[0]} rows, expected {nbins}")
            self._cols[k] = hold(v, v.dtype) if pin and v.dtype == np.float64 else v

    def _bin_columns(self):
        return list(self._cols)

    def _bin_column(self, name):
        return self._cols[name]

    def add_bin_column(self, name, values):
        self._cols[name] = np.asarray(values)

    def _chrom_of(self, lo, hi):
        ci = int(np.searchsorted(self._chrom_offset, lo, side="right") - 1)
        if hi > int(self._chrom_offset[ci + 1]):
            raise ValueError("bin range spans chromosomes")
        return ci

    def region_upper_csr(self, lo, hi):
        ci = self._chrom_of(lo, hi)
        off = int(self._chrom_offset[ci])
        name = self.chromnames[ci]
        ip, col, cnt = self._regions[name]
        if lo == off and hi == int(self._chrom_offset[ci + 1]):
            if self._lazy_pin and name not in self._pinned:
                # pin on first use: a rank of a multi-GPU job only pays for the chromosomes it actually reads
                import torch

                ts = [torch.from_numpy(a).pin_memory() for a in (ip, col, cnt)]
                self._keep.extend(ts)
                ip, col, cnt = self._regions[name] = tuple(t.numpy() for t in ts)
                self._pinned.add(name)
            return ip, col, cnt  # the stored arrays themselves
        a, b = lo - off, hi - off
        p0, p1 = int(ip[a]), int(ip[b])
        return (ip[a : b + 1] - p0).astype(np.int32), (col[p0:p1] - a).astype(np.int32), np.ascontiguousarray(cnt[p0:p1])

    def _upper_pixels(self, lo, hi):
        ci = self._chrom_of(lo, hi)
        off = int(self._chrom_offset[ci])
        ip, col, cnt = self._regions[self.chromnames[ci]]
        a, b = lo - off, hi - off
        p0, p1 = int(ip[a]), int(ip[b])
        rows = np.repeat(np.arange(a, b, dtype=np.int64), np.diff(ip[a : b + 1].astype(np.int64)))
        b2 = col[p0:p1].astype(np.int64)
        keep = b2 < b
        return rows[keep] + off, b2[keep] + off, cnt[p0:p1][keep]

    def region_csr(self, lo, hi):
        b1, b2, c = self._upper_pixels(lo, hi)
        nb = hi - lo
        offd = b1 != b2
        row = np.concatenate([b1, b2[offd]]) - lo
        col = np.concatenate([b2, b1[offd]]) - lo
        val = np.concatenate([c, c[offd]])
        order = np.lexsort((col, row))
        indptr = np.zeros(nb + 1, dtype=np.int64)
        np.cumsum(np.bincount(row, minlength=nb), out=indptr[1:])
        return indptr.astype(np.int32), col[order].astype(np.int32), val[order].astype(np.int32)

    @property
    def _bin1(self):
        raise NotImplementedError("ChromCooler keeps per-chromosome CSR arrays, not a global pixel table")

    _bin2 = _count = _bin1


def is_cooler(obj):
    """True for our cooler classes and for a real ``cooler.Cooler``."""
    if isinstance(obj, _CoolerBase):
        return True
    return type(obj).__name__ == "Cooler" and hasattr(obj, "binsize") and hasattr(obj, "matrix")


def abspath_of(clr):
    return os.path.abspath(clr.filename)

"""Cis expected-by-distance tables (input preparation, not on the pile-up hot path).

The reference consumes an *expected* DataFrame produced by
``cooltools expected-cis`` (``CLI.py:484-508``; consumed at
``coolpup.py:861-918``).  cooltools is not importable here, so this module
computes a table with the same schema (``region1, region2, dist, n_valid,
count.sum, count.avg[, balanced.sum, balanced.avg]``) for a view of cis
regions, so that expected-normalised pile-ups can be configured without any
external tool.

Semantics (cooltools ``expected_cis`` with ``smooth=False``): for every view
region and diagonal ``d``: ``n_valid`` = number of pixel positions on the
diagonal whose two bins are both valid (non-NaN weight when balanced, all bins
otherwise), ``balanced.sum`` = sum of the balanced values over those
positions, ``count.sum`` = sum of the raw counts over *every* stored pixel of
the diagonal (cooltools does not mask raw counts), ``*.avg = *.sum / n_valid``;
the first ``ignore_diags`` diagonals are NaN.  Pinned against the
cooltools-made table the reference ships (``tests/data/CN.mm9.toy_expected.tsv``)
by ``tests/test_expected.py``.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def expected_cis(clr, view_df=None, clr_weight_name=None, ignore_diags=2):
    if view_df is None:
        names = list(clr.chromnames)
        view_df = pd.DataFrame(
            {"chrom": names, "start": 0, "end": [int(clr.chromsizes[c]) for c in names], "name": names}
        )
    rows = []
    for chrom, start, end, name in zip(view_df["chrom"], view_df["start"], view_df["end"], view_df["name"]):
        lo, hi = clr.extent((chrom, start, end))
        nb = hi - lo
        b1, b2, cnt = clr._upper_pixels(lo, hi)
        d = (b2 - b1).astype(np.int64)
        dist = np.arange(nb)
        tab = {"region1": name, "region2": name, "dist": dist}
        if clr_weight_name:
            w = np.asarray(clr._bin_column(clr_weight_name), dtype=np.float64)[lo:hi]
            ok = ~np.isnan(w)
            # number of valid positions on diagonal d = sum_i ok[i] & ok[i+d]
            okf = ok.astype(np.float64)
            nfft = 1 << int(np.ceil(np.log2(max(2 * nb, 2))))
            f = np.fft.rfft(okf, nfft)
            n_valid = np.rint(np.fft.irfft(f * np.conj(f), nfft)[:nb]).astype(np.int64)
            val = w[b1 - lo] * w[b2 - lo] * cnt
            good = ~np.isnan(val)
            bal_sum = np.bincount(d[good], weights=val[good], minlength=nb)[:nb]
            # cooltools sums the raw counts over ALL stored pixels of the diagonal (masked bins included) and still
            # divides by n_valid (pinned by tests/fixtures/CN.mm9.toy_expected.tsv)
            cnt_sum = np.bincount(d, weights=cnt.astype(np.float64), minlength=nb)[:nb]
        else:
            n_valid = (nb - dist).astype(np.int64)
            cnt_sum = np.bincount(d, weights=cnt.astype(np.float64), minlength=nb)[:nb]
            bal_sum = None
        with np.errstate(divide="ignore", invalid="ignore"):
            tab["n_valid"] = n_valid
            cs = cnt_sum.copy()
            cs[:ignore_diags] = np.nan
            tab["count.sum"] = cs
            tab["count.avg"] = cs / n_valid
            if bal_sum is not None:
                bs = bal_sum.copy()
                bs[:ignore_diags] = np.nan
                tab["balanced.sum"] = bs
                tab["balanced.avg"] = bs / n_valid
        rows.append(pd.DataFrame(tab))
    return pd.concat(rows, ignore_index=True)


def expected_cis_gpu(clr, view_df=None, clr_weight_name=None, ignore_diags=2, device=0):
    """Same table as :func:`expected_cis`, with the per-diagonal sums computed on the GPU (``pup_expected_cis``) from
    the cooler's upper triangle as stored (needs a reader with ``region_upper_csr``, e.g. ``coolio.Cooler``)."""
    from . import _native

    _native.require_device()
    if view_df is None:
        names = list(clr.chromnames)
        view_df = pd.DataFrame(
            {"chrom": names, "start": 0, "end": [int(clr.chromsizes[c]) for c in names], "name": names}
        )
    rows = []
    for chrom, start, end, name in zip(view_df["chrom"], view_df["start"], view_df["end"], view_df["name"]):
        lo, hi = clr.extent((chrom, start, end))
        nb = hi - lo
        indptr, col, cnt = clr.region_upper_csr(lo, hi)
        w = None
        if clr_weight_name:
            w = np.ascontiguousarray(np.asarray(clr._bin_column(clr_weight_name), dtype=np.float64)[lo:hi])
        cs, bs, nv = _native.expected_cis_sums(device, nb, indptr, col, cnt, w)
        tab = {"region1": name, "region2": name, "dist": np.arange(nb), "n_valid": nv}
        with np.errstate(divide="ignore", invalid="ignore"):
            cs[:ignore_diags] = np.nan
            tab["count.sum"] = cs
            tab["count.avg"] = cs / nv
            if bs is not None:
                bs[:ignore_diags] = np.nan
                tab["balanced.sum"] = bs
                tab["balanced.avg"] = bs / nv
        rows.append(pd.DataFrame(tab))
    return pd.concat(rows, ignore_index=True)


def coverage(clr, ignore_diags=2):
    """Per-bin raw coverage ``(cov_cis_raw, cov_tot_raw)`` of a cooler (int64 arrays over all bins).

    Restates ``cooltools.api.coverage.coverage(clr, ignore_diags=..., store=True)``, which the reference runs when
    ``coverage_norm`` is requested and the columns are missing (``coolpup.py:955-963``): pixels closer to the diagonal
    than ``ignore_diags`` bins are zeroed; every remaining pixel adds its count to both of its bins (``cov_tot_raw``),
    cis pixels also to ``cov_cis_raw``.  (Here the columns are kept in memory instead of being written into the file.)
    """
    b1 = np.asarray(clr._bin1, dtype=np.int64)
    b2 = np.asarray(clr._bin2, dtype=np.int64)
    w = np.asarray(clr._count, dtype=np.float64).copy()
    nbins = int(clr._chrom_offset[-1])
    if ignore_diags:
        w[np.abs(b1 - b2) < ignore_diags] = 0
    chrom_of = np.searchsorted(clr._chrom_offset, np.arange(nbins), side="right") - 1
    cis = chrom_of[b1] == chrom_of[b2]
    tot = np.bincount(b1, weights=w, minlength=nbins) + np.bincount(b2, weights=w, minlength=nbins)
    cisw = w * cis
    cisc = np.bincount(b1, weights=cisw, minlength=nbins) + np.bincount(b2, weights=cisw, minlength=nbins)
    return cisc[:nbins].astype(np.int64), tot[:nbins].astype(np.int64)

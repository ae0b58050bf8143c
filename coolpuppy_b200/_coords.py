"""Vectorised window-coordinate generation (host side of the hot path).

Replaces the reference's per-row generator ``CoordCreator.pos_stream`` ->
``to_dict(orient="records")`` (``coolpup.py:598-746``) with numpy array
construction: one :class:`RegionWindows` per view region holding the first
row / first column bin, kind (ROI / control), flip flag and integer group
codes of every window, in the reference's emission order.  The random control
shifts are drawn with the same ``np.random`` calls, sizes and order as the
reference (``coolpup.py:387-453``: ``randint`` then ``choice``, once per region
for bedpe / local, once per pair offset for bed pairs), so a seeded run sees
identical windows.
"""
from __future__ import annotations

import re

import numpy as np
import pandas as pd

_tok = re.compile(r"(\d+)")


def natsorted(seq):
    """Natural sort of chromosome names (the reference uses ``natsort.natsorted``, coolpup.py:350-355, 927)."""

    def key(s):
        return tuple((0, int(p)) if p.isdigit() else (1, p) for p in _tok.split(str(s)) if p != "")

    return sorted(seq, key=key)


def default_band_edges():
    return np.append([0], 50000 * 2 ** np.arange(30))


class RegionWindows:
    """Windows of one view region, in emission order.

    ``st1``/``st2``: chromosome-relative first bins (int64); ``kind``: 0 ROI,
    1 control; ``idx1``/``idx2``: rows of ``sel`` (the region's feature table)
    on side 1 / side 2; ``distance``: centre2 - centre1 in bp (None for local).
    """

    __slots__ = ("region", "sel", "sel2", "st1", "st2", "kind", "idx1", "idx2", "distance", "paired", "frame")

    def __init__(self, region, sel, st1, st2, kind, idx1, idx2, distance, paired, sel2=None):
        self.region = region
        self.sel = sel
        self.sel2 = sel if sel2 is None else sel2  # feature table of side 2 (differs from side 1 for trans pairs)
        self.st1 = st1
        self.st2 = st2
        self.kind = kind
        self.idx1 = idx1
        self.idx2 = idx2
        self.distance = distance
        self.paired = paired  # True: bed features paired up (columns get suffix 1/2); False: bedpe rows
        self.frame = None  # full DataFrame, only materialised for user callbacks

    def __len__(self):
        return int(self.st1.shape[0])

    def take(self, sel):
        """The windows ``sel`` (index array, emission order kept) as a new RegionWindows over the same feature table."""
        if self.frame is not None:
            raise NotImplementedError("cannot subset windows after a user callback has materialised the frame")
        return RegionWindows(self.region, self.sel, self.st1[sel], self.st2[sel], self.kind[sel], self.idx1[sel],
                             self.idx2[sel], None if self.distance is None else self.distance[sel], self.paired,
                             sel2=self.sel2)

    def sizes(self):
        """(rows, columns) of every window in bins: 2 * pad + 1 everywhere, or the features' own expanded sizes for
        rescaled pile-ups (the random control shifts move a window, they do not resize it)."""
        if self.paired:
            s1 = (self.sel["endBin"].values - self.sel["stBin"].values).astype(np.int64)
            s2 = (self.sel2["endBin"].values - self.sel2["stBin"].values).astype(np.int64)
            return s1[self.idx1], s2[self.idx2]
        s1 = (self.sel["endBin1"].values - self.sel["stBin1"].values).astype(np.int64)
        s2 = (self.sel["endBin2"].values - self.sel["stBin2"].values).astype(np.int64)
        return s1[self.idx1], s2[self.idx1]

    def column(self, name, swap=None):
        """Values of 2-D interval column ``name`` for every window.

        ``swap`` (bool array) exchanges side 1 and side 2 for the flagged
        windows -- the group swap of ``flip_snip_func`` (coolpup.py:131-144).
        """
        if self.frame is not None:
            vals = self.frame[name].to_numpy()
            if swap is not None and swap.any() and name[-1] in "12":
                other = name[:-1] + ("2" if name[-1] == "1" else "1")
                if other in self.frame.columns:
                    vals = np.where(swap, self.frame[other].to_numpy(), vals)
            return vals
        if name == "distance":
            return self.distance
        if self.paired:
            if name[-1] not in "12" or name[:-1] not in self.sel.columns:
                raise KeyError(f"no 2-D interval column {name!r}")
            own = self.sel[name[:-1]].to_numpy()[self.idx1] if name[-1] == "1" else self.sel2[name[:-1]].to_numpy()[self.idx2]
            if swap is not None and swap.any():
                other = self.sel2[name[:-1]].to_numpy()[self.idx2] if name[-1] == "1" else self.sel[name[:-1]].to_numpy()[self.idx1]
                return np.where(swap, other, own)
            return own
        if name not in self.sel.columns:
            raise KeyError(f"no 2-D interval column {name!r}")
        vals = self.sel[name].to_numpy()[self.idx1]
        if swap is not None and swap.any() and name[-1] in "12":
            other = name[:-1] + ("2" if name[-1] == "1" else "1")
            if other in self.sel.columns:
                vals = np.where(swap, self.sel[other].to_numpy()[self.idx1], vals)
        return vals

    def has_column(self, name):
        if self.frame is not None:
            return name in self.frame.columns
        if name == "distance":
            return self.distance is not None
        if self.paired:
            return name[-1] in "12" and name[:-1] in self.sel.columns
        return name in self.sel.columns

    def to_frame(self):
        """The reference's 2-D interval DataFrame for this region (all columns) -- slow path for callbacks."""
        if self.frame is not None:
            return self.frame
        if self.paired:
            left = self.sel.iloc[self.idx1].reset_index(drop=True).rename(columns=lambda c: c + "1")
            right = self.sel2.iloc[self.idx2].reset_index(drop=True).rename(columns=lambda c: c + "2")
            fr = pd.concat([left, right], axis=1)
            if self.distance is not None:
                fr["distance"] = self.distance
        else:
            fr = self.sel.iloc[self.idx1].reset_index(drop=True)
        w = fr["endBin1"].values - fr["stBin1"].values
        fr["stBin1"] = self.st1
        fr["endBin1"] = self.st1 + w
        w2 = fr["endBin2"].values - fr["stBin2"].values
        fr["stBin2"] = self.st2
        fr["endBin2"] = self.st2 + w2
        fr["kind"] = np.where(self.kind == 0, "ROI", "control")
        return fr


def _draw_shifts(n, minshift, maxshift, resolution):
    """One control-shift draw of the reference (coolpup.py:392-396, 442-445): bins to add to all four bin columns."""
    shift = np.random.randint(minshift, maxshift, n)
    sign = np.random.choice([-1, 1], n)
    shift = shift * sign
    return np.round(shift / resolution).astype(np.int64)


def _chrom_is(cc, col, chrom):
    """``cc.intervals[col] == chrom`` through integer codes made once per feature table (a string comparison of the
    whole column per view region is the host path's largest cost for 1e5+ features)."""
    df = cc.intervals
    cache = cc.__dict__.setdefault("_chrom_code_cache", {})
    key = (col, id(df), len(df))
    if key not in cache:
        codes, uniques = pd.factorize(df[col])
        cache[key] = (np.asarray(codes), {str(u): i for i, u in enumerate(uniques)})
    codes, lut = cache[key]
    return codes == lut.get(str(chrom), -2)


def build_region_windows(cc, region, control, draw_only=False):
    """Windows of one view region ``(chrom, start, end)`` (reference: coolpup.py:546-563, 598-746).

    ``draw_only``: make exactly the ``np.random`` calls the region's control shifts need (same sizes, same order) and
    return ``None`` -- how a rank keeps the global random stream in step with the serial reference for view regions
    that another rank piles up, without laying out their windows."""
    chrom, start, end = region
    df = cc.intervals
    nctrl = cc.nshifts if control else 0
    res = cc.resolution
    if cc.kind == "bedpe":
        m = (
            _chrom_is(cc, "chrom1", chrom) & _chrom_is(cc, "chrom2", chrom)
            & (df["start1"].values >= start) & (df["end1"].values < end)
            & (df["start2"].values >= start) & (df["end2"].values < end)
        )
        sel = df[m].reset_index(drop=True)
        q = len(sel)
        idx = np.arange(q, dtype=np.int64)
        st1 = sel["stBin1"].values.astype(np.int64)
        st2 = sel["stBin2"].values.astype(np.int64)
        dist = sel["distance"].values.astype(np.float64)
        kind = np.zeros(q, dtype=np.int8)
        if nctrl > 0 and q > 0:
            dbin = _draw_shifts(q * nctrl, cc.minshift, cc.maxshift, res)
            if draw_only:
                return None
            cidx = np.tile(idx, nctrl)
            st1 = np.concatenate([st1, st1[cidx] + dbin])
            st2 = np.concatenate([st2, st2[cidx] + dbin])
            dist = np.concatenate([dist, dist[cidx]])
            kind = np.concatenate([kind, np.ones(q * nctrl, dtype=np.int8)])
            idx = np.concatenate([idx, cidx])
        return RegionWindows(region, sel, st1, st2, kind, idx, idx, dist, paired=False)

    m = _chrom_is(cc, "chrom", chrom) & (df["start"].values >= start) & (df["end"].values < end)
    sel = df[m].reset_index(drop=True)
    nfeat = len(sel)
    stbin = sel["stBin"].values.astype(np.int64)
    if cc.local:
        idx = np.arange(nfeat, dtype=np.int64)
        st = stbin.copy()
        kind = np.zeros(nfeat, dtype=np.int8)
        st1 = st2 = st
        if nctrl > 0 and nfeat > 0:
            dbin = _draw_shifts(nfeat * nctrl, cc.minshift, cc.maxshift, res)
            if draw_only:
                return None
            cidx = np.tile(idx, nctrl)
            st1 = st2 = np.concatenate([st, st[cidx] + dbin])
            kind = np.concatenate([kind, np.ones(nfeat * nctrl, dtype=np.int8)])
            idx = np.concatenate([idx, cidx])
        return RegionWindows(region, sel, st1, st2, kind, idx, idx, None, paired=True)

    # all ordered pairs (k, k + i), by offset i then k, distance-filtered; per offset block the ROI rows, then the
    # nshifts shifted replicas.  The layout is native host code (pup_pair_windows_*); only the np.random calls, whose
    # sizes and order pin the reference's stream (coolpup.py:392-396, 697-699: one draw per offset), stay here.
    from . import _native

    center = sel["center"].values.astype(np.float64)
    q, total = _native.pair_windows_count(center, cc.mindist, cc.maxdist)
    dbin = None
    if draw_only:
        for n in q[q > 0] * nctrl:
            np.random.randint(cc.minshift, cc.maxshift, int(n))
            np.random.choice([-1, 1], int(n))
        return None
    if nctrl > 0 and total > 0:
        shift = np.empty(total * nctrl, dtype=np.int64)
        sign = np.empty(total * nctrl, dtype=np.int64)
        pos = 0
        for n in q[q > 0] * nctrl:
            n = int(n)
            shift[pos : pos + n] = np.random.randint(cc.minshift, cc.maxshift, n)
            sign[pos : pos + n] = np.random.choice([-1, 1], n)
            pos += n
        dbin = np.round(shift * sign / res).astype(np.int64)
    else:
        nctrl = 0
    st1, st2, kind, kk, ll, dist_all = _native.pair_windows_fill(stbin, center, cc.mindist, cc.maxdist, nctrl, dbin, total)
    return RegionWindows(region, sel, st1, st2, kind, kk, ll, dist_all, paired=True)


def _draw_shifts_trans(n, minshift, maxshift, resolution):
    """Control-shift draw of a TRANS block (coolpup.py:392-407): a second (shift2, sign2) pair is drawn for side 2, but
    the bin columns of BOTH sides move by the first shift (430-433) -- the second draw only advances the stream."""
    dbin = _draw_shifts(n, minshift, maxshift, resolution)
    np.random.randint(minshift, maxshift, n)
    np.random.choice([-1, 1], n)
    return dbin


def build_trans_windows(cc, region1, region2, control, draw_only=False):
    """Windows between two view regions on different chromosomes (reference: coolpup.py:565-590, 652-680, 1330-1348).

    bed: every feature of region1 with every feature of region2 (``itertools.product``), one ``_control_regions`` call
    PER PAIR (ROI row, then its nshifts controls); bedpe: the rows joining the two regions in either orientation -- the
    rows stored the other way round are taken as they are, their side-1 bins applied to region1 (a quirk of
    ``_filter_func_trans_pairs``, reproduced) -- and one ``_control_regions`` call for the region pair."""
    df = cc.intervals
    nctrl = cc.nshifts if control else 0
    res = cc.resolution
    (ch1, s1, e1), (ch2, s2, e2) = region1, region2
    if cc.kind == "bedpe":
        a = ((df["chrom1"].values == ch1) & (df["chrom2"].values == ch2) & (df["start1"].values >= s1) & (df["end1"].values < e1)
             & (df["start2"].values >= s2) & (df["end2"].values < e2))
        b = ((df["chrom2"].values == ch1) & (df["chrom1"].values == ch2) & (df["start2"].values >= s1) & (df["end2"].values < e1)
             & (df["start1"].values >= s2) & (df["end1"].values < e2))
        sel = pd.concat([df[a], df[b]]).reset_index(drop=True)
        q = len(sel)
        idx = np.arange(q, dtype=np.int64)
        st1 = sel["stBin1"].values.astype(np.int64)
        st2 = sel["stBin2"].values.astype(np.int64)
        dist = sel["distance"].values.astype(np.float64)
        kind = np.zeros(q, dtype=np.int8)
        if nctrl > 0 and q > 0:
            dbin = _draw_shifts_trans(q * nctrl, cc.minshift, cc.maxshift, res)
            if draw_only:
                return None
            cidx = np.tile(idx, nctrl)
            st1 = np.concatenate([st1, st1[cidx] + dbin])
            st2 = np.concatenate([st2, st2[cidx] + dbin])
            dist = np.concatenate([dist, dist[cidx]])
            kind = np.concatenate([kind, np.ones(q * nctrl, dtype=np.int8)])
            idx = np.concatenate([idx, cidx])
        return RegionWindows((region1, region2), sel, st1, st2, kind, idx, idx, dist, paired=False)
    left = df[(df["chrom"].values == ch1) & (df["start"].values >= s1) & (df["end"].values < e1)].reset_index(drop=True)
    right = df[(df["chrom"].values == ch2) & (df["start"].values >= s2) & (df["end"].values < e2)].reset_index(drop=True)
    n1, n2 = len(left), len(right)
    x = np.repeat(np.arange(n1, dtype=np.int64), n2)
    y = np.tile(np.arange(n2, dtype=np.int64), n1)
    lb = left["stBin"].values.astype(np.int64)
    rb = right["stBin"].values.astype(np.int64)
    if nctrl == 0 or n1 * n2 == 0:
        return RegionWindows((region1, region2), left, lb[x], rb[y], np.zeros(n1 * n2, dtype=np.int8), x, y, None,
                             paired=True, sel2=right)
    shifts = np.empty((n1 * n2, nctrl), dtype=np.int64)
    for p in range(n1 * n2):  # one draw per pair: that is the reference's stream
        shifts[p] = _draw_shifts_trans(nctrl, cc.minshift, cc.maxshift, res)
    if draw_only:
        return None
    block = 1 + nctrl  # per pair: the ROI row, then its controls
    st1 = np.repeat(lb[x], block)
    st2 = np.repeat(rb[y], block)
    add = np.concatenate([np.zeros((n1 * n2, 1), dtype=np.int64), shifts], axis=1).reshape(-1)
    kind = np.tile(np.concatenate([[0], np.ones(nctrl, dtype=np.int8)]).astype(np.int8), n1 * n2)
    return RegionWindows((region1, region2), left, st1 + add, st2 + add, kind, np.repeat(x, block), np.repeat(y, block),
                         None, paired=True, sel2=right)

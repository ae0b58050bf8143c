"""CPU ORACLE for the pile-up hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This is a plain numpy/scipy restatement of the reference algorithm
(open2c/coolpuppy 1.1.0 @592673c), written from the reference's behaviour and
citing the lines it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product package ``coolpuppy_b200`` never does.

Parity pinning: the unmodified reference itself runs in the build container
through ``oracle/refshim`` (stand-ins for its absent third-party imports) and
``tests/golden/make_golden.py`` stores its outputs -- final pile-ups, the
per-region window streams (incl. ``np.random`` control-shift order) and the
per-region accumulators -- under ``tests/golden/``.  ``tests/test_oracle.py``
checks this restatement against every one of those vectors, against the
reference-asserted known answers (``tests/test_coolpup.py:50-72, 97, 124-142,
167-172``) and statistically against the legacy ``loop_ref.np.txt``.  What
remains unpinned by the reference's *own* tests: the third-party pieces
(cooler's balanced fetch, cooltools' LazyToeplitz / ExpectedSnipper), which are
restated from their documentation in ``oracle/refshim`` and
``coolpuppy_b200/coolio.py``.

Scope: cis pile-ups (bed all-pairs, bed local, bedpe), random-shift controls,
expected (ooe and not), balancing weights, coverage normalisation, grouping by
strand / distance / arbitrary columns / window, flips, stripes.  Not restated:
``trans`` and ``rescale`` (out of scope).
"""
from __future__ import annotations

import re
from functools import reduce

import numpy as np
import pandas as pd
from scipy import sparse

__all__ = ["oracle_pileup", "OracleResult", "key_repr"]


# --------------------------------------------------------------------------- helpers
def natsorted(seq):
    """natsort.natsorted for chromosome names (coolpup.py:350-355, 927)."""
    tok = re.compile(r"(\d+)")

    def key(s):
        return tuple((0, int(p)) if p.isdigit() else (1, p) for p in tok.split(str(s)) if p != "")

    return sorted(seq, key=key)


def key_repr(g):
    """Stable text form of a group key (same convention as tests/golden/make_golden.py)."""
    if isinstance(g, str):
        return g
    return repr(tuple(x.item() if isinstance(x, np.generic) else x for x in g))


def default_band_edges():
    """coolpup.py:46-47."""
    return np.append([0], 50000 * 2 ** np.arange(30))


# --------------------------------------------------------------------------- a1: windows
def expand_windows(center, flank, resolution):
    """[exp_start, exp_end) of a feature centre (coolpup.py:78-86, 94-107).

    ``exp_start = floor(center / res) * res - flank``,
    ``exp_end   = floor(center / res + 1) * res + flank``.
    """
    center = np.asarray(center, dtype=float)
    exp_start = np.floor(center / resolution) * resolution - flank
    exp_end = np.floor(center / resolution + 1) * resolution + flank
    return exp_start, exp_end


def expand_scale(start, end, scale):
    """``bioframe.expand(df, scale=scale)`` as called by coolpup.py:87-90, 108-114 (third-party, restated from its source
    as remembered: grow about the midpoint to ``scale`` times the length, ``DataFrame.round`` = half to even, cast back to
    the integer dtype).  No artefact in the reference tree pins it."""
    start, end = np.asarray(start), np.asarray(end)
    pads = 0.5 * (scale - 1) * (end - start)
    return np.round(start - pads).astype(start.dtype), np.round(end + pads).astype(end.dtype)


def zoom_array(in_array, final_shape):
    """``cooltools.lib.numutils.zoom_array`` (third-party; called by coolpup.py:1223-1233), restated from its source as
    remembered; ``scipy.ndimage.zoom`` is the installed scipy."""
    from scipy.ndimage import zoom

    in_array = np.asarray(in_array, dtype=np.double)
    mults = [int(np.ceil(i / f)) if f < i else 1 for i, f in zip(in_array.shape, final_shape)]
    temp_shape = tuple(f * m for f, m in zip(final_shape, mults))
    rescaled = zoom(in_array, np.array(temp_shape) / np.array(in_array.shape) + 0.0000001, order=1)
    for ind, mult in enumerate(mults):
        if mult != 1:
            sh = list(rescaled.shape)
            rescaled.shape = sh[:ind] + [sh[ind] // mult, mult] + sh[ind + 1 :]
            rescaled = np.mean(rescaled, axis=ind + 1)
    return rescaled


def rescale_snip(snip, rescale_size, local, coverage_norm):
    """``PileUpper._rescale_snip`` (coolpup.py:1193-1234)."""
    import warnings

    data = snip["data"]
    if data.size == 0 or np.all(np.isnan(data)):
        snip["data"] = np.zeros((rescale_size, rescale_size))
    else:
        if local:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", category=RuntimeWarning)
                data = np.nanmean(np.dstack((data, data.T)), 2)
        nans = np.isnan(data) * 1
        data = zoom_array(np.nan_to_num(data), (rescale_size, rescale_size))
        nanzoom = zoom_array(nans, (rescale_size, rescale_size))
        data[np.ceil(nanzoom).astype(bool)] = np.nan
        with np.errstate(divide="ignore", invalid="ignore"):
            data = data * (1 / np.isfinite(nanzoom))
        snip["data"] = data
    if coverage_norm:
        snip["cov_start"] = zoom_array(snip["cov_start"], (rescale_size,))
        snip["cov_end"] = zoom_array(snip["cov_end"], (rescale_size,))
    return snip


def to_bins(exp_start, exp_end, resolution):
    """stBin = floor(exp_start/res), endBin = ceil(exp_end/res) (coolpup.py:492-497, 503-514)."""
    return (
        np.floor(np.asarray(exp_start) / resolution).astype(int),
        np.ceil(np.asarray(exp_end) / resolution).astype(int),
    )


class Coords:
    """Restatement of ``CoordCreator.process`` (coolpup.py:259-385)."""

    def __init__(self, features, resolution, *, features_format="auto", flank=100000, chroms="all",
                 minshift=10**5, maxshift=10**6, nshifts=10, mindist="auto", maxdist=None, local=False,
                 subset=0, seed=None, trans=False, rescale_flank=None):
        df = features.copy()
        self.resolution = resolution
        self.flank = flank
        self.minshift, self.maxshift, self.nshifts = minshift, maxshift, nshifts
        self.trans = trans
        if isinstance(mindist, str) and mindist == "auto":  # 240-253: explicit distances are ignored for trans
            self.mindist = 2 * flank + 2 * resolution
        else:
            self.mindist = 0 if trans else mindist
        self.maxdist = np.inf if (maxdist is None or trans) else maxdist
        self.local = local
        if features_format in (None, "auto"):  # 260-277
            if all(c in df.columns for c in ["chrom1", "start1", "end1", "chrom2", "start2", "end2"]):
                self.kind = "bedpe"
            elif all(c in df.columns for c in ["chrom", "start", "end"]):
                self.kind = "bed"
            else:
                raise ValueError("cannot determine feature kind")
        else:
            self.kind = features_format
        if subset > 0:  # 281-282, 455-461
            if seed is not None:
                np.random.seed(seed)
            if subset < len(df):
                df = df.sample(subset)
        if self.kind == "bed":  # 284-294
            df["chrom"] = df["chrom"].astype(str)
            df["center"] = (df["start"] + df["end"]) / 2
            if rescale_flank is None:
                df["exp_start"], df["exp_end"] = expand_windows(df["center"], flank, resolution)
            else:
                df["exp_start"], df["exp_end"] = expand_scale(df["start"].values, df["end"].values, 2 * rescale_flank + 1)
        else:  # 295-321
            df[["chrom1", "chrom2"]] = df[["chrom1", "chrom2"]].astype(str)
            df["center1"] = (df["start1"] + df["end1"]) / 2
            df["center2"] = (df["start2"] + df["end2"]) / 2
            df["distance"] = df["center2"] - df["center1"]
            df = df[(self.mindist <= df["distance"].abs()) & (df["distance"].abs() <= self.maxdist)]
            df = df.reset_index(drop=True)
            for sd in ("1", "2"):
                if rescale_flank is None:
                    df["exp_start" + sd], df["exp_end" + sd] = expand_windows(df["center" + sd], flank, resolution)
                else:
                    df["exp_start" + sd], df["exp_end" + sd] = expand_scale(df["start" + sd].values, df["end" + sd].values,
                                                                            2 * rescale_flank + 1)
        self.empty = df.shape[0] == 0  # 323-331
        if self.empty:
            self.final_chroms = []
            self.intervals = df
            return
        if self.kind == "bed":  # 336-356
            base = set(df["chrom"])
        else:
            if local:
                raise ValueError("Can't make local with both sides of loops defined")
            if trans:  # 341-345
                base = set(df["chrom1"].unique().tolist() + df["chrom2"].unique().tolist())
            else:
                base = set(df["chrom1"]).intersection(set(df["chrom2"]))
        self.final_chroms = natsorted(base) if chroms == "all" else natsorted(set(chroms) & base)
        if not self.final_chroms:
            raise ValueError("No chromosomes are in common between the coordinate file and the cooler file")
        if self.kind == "bed":  # 489-500
            df = df.sort_values(["chrom", "start"])
            df["stBin"], df["endBin"] = to_bins(df["exp_start"], df["exp_end"], resolution)
        else:  # 501-520
            df = df.sort_values(["chrom1", "chrom2", "start1", "start2"])
            df["stBin1"], df["endBin1"] = to_bins(df["exp_start1"], df["exp_end1"], resolution)
            df["stBin2"], df["endBin2"] = to_bins(df["exp_start2"], df["exp_end2"], resolution)
        self.intervals = df

    # ------------------------------------------------------------------ a2: control shifts
    def control_shifts(self, frame, nshifts):
        """Random-shift controls appended after the ROI rows (coolpup.py:387-453).

        ``frame`` rows are replicated ``nshifts`` times block-wise; one
        ``np.random.randint(minshift, maxshift)`` and one
        ``np.random.choice([-1, 1])`` per replicated row (drawn as two vectors,
        in that order); the same signed shift moves both sides by
        ``np.round(shift / res)`` bins (half-to-even).
        """
        frame = frame.copy()
        frame["kind"] = "ROI"
        if nshifts <= 0:
            return frame
        ctrl = pd.concat([frame] * nshifts).reset_index(drop=True)
        shift = np.random.randint(self.minshift, self.maxshift, ctrl.shape[0])
        sign = np.random.choice([-1, 1], ctrl.shape[0])
        shift = shift * sign
        if self.trans:  # 397-407: a second shift is drawn for side 2, but the BIN columns all move by the first one (430-433)
            np.random.randint(self.minshift, self.maxshift, ctrl.shape[0])
            np.random.choice([-1, 1], ctrl.shape[0])
        dbin = np.round(shift / self.resolution).astype(int)
        for c in ("stBin1", "endBin1", "stBin2", "endBin2"):
            ctrl[c] = ctrl[c].values + dbin
        for c in ("center1", "center2"):
            ctrl[c] = ctrl[c].values + shift
        ctrl["kind"] = "control"
        return pd.concat([frame, ctrl]).reset_index(drop=True)

    # ------------------------------------------------------------------ a2: per-region streams
    def region_frames(self, region, control, modify=None):
        """Frames of 2-D intervals of one view region, in the reference's emission order.

        bedpe: ``get_intervals_stream`` (coolpup.py:716-746) with the region
        filter of 554-563.  bed: ``get_combinations`` (598-714) with the filter
        of 546-552 -- ``local`` pairs every feature with itself (620-627), else
        all ordered pairs by offset ``i`` then position ``k`` (682-699) with
        controls drawn once per offset.
        """
        chrom, start, end = region
        df = self.intervals
        if self.kind == "bedpe":
            sel = df[(df["chrom1"] == chrom) & (df["chrom2"] == chrom) & (df["start1"] >= start) & (df["end1"] < end)
                     & (df["start2"] >= start) & (df["end2"] < end)].reset_index(drop=True)
            fr = self.control_shifts(sel, self.nshifts * control)
            if modify is not None:
                fr = modify(fr)
            if len(fr):
                yield fr
            return
        sel = df[(df["chrom"] == chrom) & (df["start"] >= start) & (df["end"] < end)].reset_index(drop=True)
        left = sel.rename(columns=lambda c: c + "1")
        right = sel.rename(columns=lambda c: c + "2")
        if self.local:
            merged = pd.concat([left, right], axis=1)
            fr = self.control_shifts(merged, self.nshifts * control)
            if modify is not None:
                fr = modify(fr)
            if len(fr):
                yield fr
            return
        m = len(sel)
        for i in range(1, m):  # offsets >= m give empty frames in the reference (682-689)
            comb = pd.concat([left.iloc[:-i].reset_index(drop=True), right.iloc[i:].reset_index(drop=True)], axis=1)
            comb["distance"] = comb["center2"] - comb["center1"]
            comb = comb[(self.mindist <= comb["distance"].abs()) & (comb["distance"].abs() <= self.maxdist)]
            fr = self.control_shifts(comb, self.nshifts * control)
            if modify is not None:
                fr = modify(fr)
            if len(fr):
                yield fr


def trans_frames(cc, region1, region2, control, modify=None):
    """Frames of the 2-D intervals between two view regions on different chromosomes (coolpup.py:565-590, 652-680,
    1330-1348): bedpe -- rows joining the regions in either stored orientation, concatenated, one control draw; bed --
    ``itertools.product`` of the two feature lists, one single-row frame (and control draw) per pair."""
    (c1, s1, e1), (c2, s2, e2) = region1, region2
    df = cc.intervals
    if cc.kind == "bedpe":
        a = df[(df["chrom1"] == c1) & (df["chrom2"] == c2) & (df["start1"] >= s1) & (df["end1"] < e1)
               & (df["start2"] >= s2) & (df["end2"] < e2)].reset_index(drop=True)
        b = df[(df["chrom2"] == c1) & (df["chrom1"] == c2) & (df["start2"] >= s1) & (df["end2"] < e1)
               & (df["start1"] >= s2) & (df["end1"] < e2)].reset_index(drop=True)
        fr = cc.control_shifts(pd.concat([a, b]), cc.nshifts * control)
        if modify is not None:
            fr = modify(fr)
        if len(fr):
            yield fr
        return
    left = df[(df["chrom"] == c1) & (df["start"] >= s1) & (df["end"] < e1)].reset_index(drop=True).rename(columns=lambda c: c + "1")
    right = df[(df["chrom"] == c2) & (df["start"] >= s2) & (df["end"] < e2)].reset_index(drop=True).rename(columns=lambda c: c + "2")
    for x in range(len(left)):
        for y in range(len(right)):
            comb = pd.concat([left.iloc[[x]].reset_index(drop=True), right.iloc[[y]].reset_index(drop=True)], axis=1)
            fr = cc.control_shifts(comb, cc.nshifts * control)
            if modify is not None:
                fr = modify(fr)
            yield fr


def band_annotator(edges):
    """``bin_distance_intervals`` (coolpup.py:28-51)."""
    edges = default_band_edges() if isinstance(edges, str) and edges == "default" else np.asarray(edges)

    def annotate(fr):
        ids = np.searchsorted(edges, fr["distance"], side="right")
        fr = fr.copy()
        fr["distance_band"] = [tuple(edges[i - 1 : i + 1]) for i in ids]
        return fr

    return annotate


# --------------------------------------------------------------------------- a6/a7: accumulate
def add_snip(store, key, snip):
    """``_add_snip`` (lib/puputils.py:12-38): first snippet kept verbatim, later ones nansum'ed."""
    if key not in store:
        store[key] = {
            "data": snip["data"],
            "cov_start": snip["cov_start"],
            "cov_end": snip["cov_end"],
            "num": np.isfinite(snip["data"]).astype(int),
            "n": 1,
            "coordinates": [snip["coordinates"]],
            "horizontal_stripe": [snip["horizontal_stripe"]],
            "vertical_stripe": [snip["vertical_stripe"]],
        }
    else:
        p = store[key]
        p["data"] = np.nansum([p["data"], snip["data"]], axis=0)
        p["num"] = p["num"] + np.isfinite(snip["data"]).astype(int)
        p["cov_start"] = np.nansum([p["cov_start"], snip["cov_start"]], axis=0)
        p["cov_end"] = np.nansum([p["cov_end"], snip["cov_end"]], axis=0)
        p["n"] += 1
        for f in ("coordinates", "horizontal_stripe", "vertical_stripe"):
            p[f] = p[f] + [snip[f]]


def sum_pups(p1, p2):
    """``sum_pups`` (lib/puputils.py:88-113): NaN -> 0 and +inf -> 1.797e308 on both inputs, then add."""
    d1 = np.nan_to_num(p1["data"])
    d2 = np.nan_to_num(p2["data"])
    return {
        "data": d1 + d2,
        "cov_start": p1["cov_start"] + p2["cov_start"],
        "cov_end": p1["cov_end"] + p2["cov_end"],
        "n": p1["n"] + p2["n"],
        "num": p1["num"] + p2["num"],
        "horizontal_stripe": p1["horizontal_stripe"] + p2["horizontal_stripe"],
        "vertical_stripe": p1["vertical_stripe"] + p2["vertical_stripe"],
        "coordinates": p1["coordinates"] + p2["coordinates"],
    }


def norm_coverage(pup):
    """``norm_coverage`` (lib/puputils.py:168-190)."""
    cov = np.outer(pup["cov_start"], pup["cov_end"])
    with np.errstate(divide="ignore", invalid="ignore"):
        cov = cov / np.nanmean(cov)
        data = pup["data"] / cov
    data[np.isnan(data)] = 0
    out = dict(pup)
    out["data"] = data
    return out


class OracleResult:
    def __init__(self, rows, regions, windows, params):
        self.rows = rows  # list of dicts in the reference's final row order
        self.regions = regions  # region -> {(kind, key_repr): pup}
        self.windows = windows  # region -> dict of arrays (st1, st2, kind, flip, group key reprs)
        self.params = params

    def by_key(self):
        return {key_repr(r["group"]): r for r in self.rows}


# --------------------------------------------------------------------------- the path
def oracle_pileup(clr, features, features_format="bed", view_df=None, expected_df=None,
                  expected_value_col="balanced.avg", clr_weight_name="weight", flank=100000, minshift=10**5,
                  maxshift=10**6, nshifts=0, ooe=True, mindist="auto", maxdist=None, min_diag=2, subset=0,
                  by_window=False, by_strand=False, by_distance=False, groupby=[], ignore_group_order=False,
                  flip_negative_strand=False, local=False, coverage_norm=False, store_stripes=False, seed=None,
                  max_windows_per_region=None, trans=False, rescale=False, rescale_flank=1, rescale_size=99):
    """Restatement of ``pileup()`` -> ``PileUpper.pileupsWith*Control`` (coolpup.py:1922-2279, 1360-1919).

    ``max_windows_per_region`` is NOT a reference feature: it truncates every
    region's ROI+control stream after that many windows so that ``bench.py``
    can time a bounded sample of a large workload on the CPU.
    """
    # ---- pileup(): argument normalisation (2088-2138)
    distance_edges = None
    if by_distance is not False:
        if local:
            raise ValueError("Can't do local pileups by distance")
        if isinstance(by_distance, np.ndarray):
            distance_edges = [int(i) for i in by_distance]
        elif by_distance is True or (isinstance(by_distance, str) and by_distance == "default"):
            distance_edges = "default"
        else:
            raise ValueError("Invalid by_distance value")
        by_distance = True
    if seed is not None:
        np.random.seed(seed)
    resolution = clr.binsize
    if view_df is None:
        names = list(clr.chromnames)
        view = pd.DataFrame({"chrom": names, "start": 0, "end": [int(clr.chromsizes[c]) for c in names], "name": names})
    else:
        view = view_df.copy()
        if "name" not in view.columns:
            view["name"] = view["chrom"]
    control = nshifts > 0
    chroms = list(view["chrom"].unique())
    if by_window and (features_format != "bed" or local):
        raise ValueError("by-window needs bed features and non-local pile-ups")

    cc = Coords(features, resolution, features_format=features_format, flank=flank, chroms=chroms, minshift=minshift,
                maxshift=maxshift, nshifts=nshifts, mindist=mindist, maxdist=maxdist, local=local, subset=subset,
                seed=seed, trans=trans, rescale_flank=rescale_flank if rescale else None)
    if rescale and rescale_size % 2 == 0:
        raise ValueError("Please provide an odd rescale_size")
    if trans and by_distance:
        raise ValueError("Cannot do by-distance pileups for trans")
    if trans and local:
        raise ValueError("Cannot do local with trans=True")

    # ---- PileUpper.__init__ (837-997)
    pad = flank // resolution
    W = rescale_size if rescale else 2 * pad + 1
    rescale_to = rescale_size if rescale else None
    expected = expected_df is not None and expected_df is not False
    E = {}
    if expected:
        ex = expected_df[expected_df["region1"].isin(view["name"]) & expected_df["region2"].isin(view["name"])]
        if control:  # 867-872
            control = False
        if trans:  # get_expected_trans (999-1005): one scalar per region pair
            for _, row in ex.iterrows():
                E[(row["region1"], row["region2"])] = float(row[expected_value_col])
        else:
            ex = ex[ex["region1"] == ex["region2"]].reset_index(drop=True)
            for name in view["name"]:  # ExpectedSnipper.select: table row order (907-916)
                E[name] = ex.loc[(ex["region1"] == name) & (ex["region2"] == name), expected_value_col].values.astype(float)
    extents = {}
    for _, r in view.iterrows():  # 922-925
        lo, hi = clr.extent((r["chrom"], r["start"], r["end"]))
        off = clr.offset(r["chrom"])
        extents[r["name"]] = (lo - off, hi - off)
    use_chroms = natsorted(set(cc.final_chroms) & set(clr.chromnames))  # 927-930
    view = view[view["chrom"].isin(use_chroms)].set_index("name")
    if coverage_norm is True:  # 944-949
        coverage_norm = "cov_tot_raw"
    elif coverage_norm == "cis":
        coverage_norm = "cov_cis_raw"
    elif coverage_norm == "total":
        coverage_norm = "cov_tot_raw"
    if coverage_norm and clr_weight_name:
        raise ValueError("Can't do coverage normalization when clr_weight_name is provided")

    # ---- wrappers (1656-1919)
    groupby = list(groupby)
    modify = None
    dup_by_region = False
    if by_window:
        dup_by_region = True
        groupby = []
    elif by_strand and by_distance:
        groupby = ["strand1", "strand2", "distance_band"] + groupby
    elif by_strand:
        groupby = ["strand1", "strand2"] + groupby
    elif by_distance:
        groupby = ["distance_band"] + groupby
    if by_distance:
        edges = distance_edges
        if not (isinstance(edges, str) and edges == "default"):  # 1789-1797
            edges = list(np.sort(edges))
            for _ in range(len(edges)):
                if np.min(edges) < cc.mindist:
                    edges[int(np.argmin(edges))] = cc.mindist
                else:
                    break
        modify = band_annotator(edges)

    # ---- flip wiring (1431-1493)
    flipby = None
    do_flip = False
    if flip_negative_strand:
        flipby = "strand"
        do_flip = True
    elif ignore_group_order and groupby:
        g = np.array(groupby)
        filt = [f"{x}1" in g and f"{x}2" in g for x in [y[:-1] for y in g]]
        gf = np.sort(g[filt])
        if ignore_group_order is True:
            fb = list(set(x[:-1] for x in gf))
        elif isinstance(ignore_group_order, str):
            fb = [ignore_group_order]
        elif len(ignore_group_order) == 1:
            fb = list(ignore_group_order)
        else:
            fb = list(set(x[:-1] for x in ignore_group_order))
        if len(fb) == 1 and f"{fb[0]}1" in gf:
            flipby = fb[0]
        else:
            raise ValueError("Ambiguous ignore_group_order")
        do_flip = True

    def modify_final(fr):  # flip_mark_intervals_func (118-125)
        if do_flip:
            fr = fr.copy()
            if flip_negative_strand:
                fr["flip"] = np.where(fr["strand1"] == "-", True, False)
            else:
                fr["flip"] = fr[f"{flipby}1"] > fr[f"{flipby}2"]
        if modify is not None:
            fr = modify(fr)
        return fr

    # ---- per region (1285-1358)
    if cc.empty or len(view) == 0:
        return OracleResult([], {}, {}, {"W": W})
    region_out = {}
    window_log = {}
    if trans:  # every pair of view regions on different chromosomes (1419-1426)
        import itertools

        for n1, n2 in itertools.combinations(view.index, 2):
            ra, rb = view.loc[n1], view.loc[n2]
            if ra["chrom"] == rb["chrom"]:
                continue
            region_out[(n1, n2)], window_log[(n1, n2)] = _pileup_region(
                clr, cc, (n1, n2), ((ra["chrom"], ra["start"], ra["end"]), (rb["chrom"], rb["start"], rb["end"])),
                (extents[n1], extents[n2]), control, modify_final, groupby, do_flip, ignore_group_order, dup_by_region,
                E.get((n1, n2)) if expected else None, ooe, clr_weight_name, coverage_norm, min_diag, W, store_stripes,
                max_windows_per_region, trans=True, rescale_to=rescale_to, local=local)
    for rname, r in ([] if trans else view.iterrows()):
        region_out[rname], window_log[rname] = _pileup_region(
            clr, cc, rname, (r["chrom"], r["start"], r["end"]), extents[rname], control, modify_final, groupby,
            do_flip, ignore_group_order, dup_by_region, E.get(rname) if expected else None, ooe, clr_weight_name,
            coverage_norm, min_diag, W, store_stripes, max_windows_per_region, rescale_to=rescale_to, local=local)

    # ---- reduce over regions, keys in order of first appearance (1511-1531)
    def reduce_kind(kind):
        order = []
        for rname in region_out:
            for k in region_out[rname][kind]:
                if k not in order:
                    order.append(k)
        return {k: reduce(sum_pups, [region_out[rn][kind][k] for rn in region_out if k in region_out[rn][kind]])
                for k in order}

    roi = reduce_kind("ROI")
    has_ctrl = control or (expected and not ooe)
    ctrl = reduce_kind("control") if has_ctrl else None
    # ---- final normalisation (1533-1607)
    if coverage_norm:
        roi = {k: norm_coverage(v) for k, v in roi.items()}
        if control:
            ctrl = {k: norm_coverage(v) for k, v in ctrl.items()}
    rows = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for k, p in roi.items():
            data = p["data"] / p["num"]
            row = {"group": k, "n": p["n"], "num": p["num"]}
            if has_ctrl:
                if k in ctrl:
                    c = ctrl[k]
                    data = data / (c["data"] / c["num"])
                    row["control_n"], row["control_num"] = c["n"], c["num"]
                else:  # pandas index alignment gives NaN (1546-1548)
                    data = data * np.nan
                    row["control_n"], row["control_num"] = np.nan, np.nan
            data = np.where(data == np.inf, np.nan, data)
            if store_stripes:
                row["coordinates"] = [x.split(".") for x in p["coordinates"]]
                hs, vs = p["horizontal_stripe"], p["vertical_stripe"]
                if has_ctrl:
                    call = ctrl["all"]
                    cn = call["data"] / call["num"]
                    cntr = int(np.floor(cn.shape[0] / 2))
                    hs = [np.divide(x, cn[cntr, :]) for x in hs]
                    vs = [np.divide(x, cn[:, cntr][::-1]) for x in vs]
                row["horizontal_stripe"] = np.vstack(hs)
                row["vertical_stripe"] = np.vstack(vs)
                if local:  # numutils._copy_array_halves (coolpup.py:1594-1600, lib/numutils.py:6-9)
                    cn = int(np.floor(row["vertical_stripe"].shape[1] / 2))
                    for f in ("vertical_stripe", "horizontal_stripe"):
                        x = row[f]
                        x[:, : cn + 1] = np.fliplr(x[:, cn:])
            if local:
                import warnings

                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", category=RuntimeWarning)
                    data = np.nanmean(np.dstack((data, data.T)), 2)
            row["data"] = data
            rows.append(row)
    if by_distance:  # 1805-1807, 1890-1892
        rows = [r for r in rows if not (isinstance(r["group"], tuple) and () in r["group"])]
    return OracleResult(rows, {rn: {(kd, key_repr(k)): p for kd in v for k, p in v[kd].items()} for rn, v in region_out.items()},
                        window_log, {"W": W, "groupby": groupby})


def _pileup_region(clr, cc, rname, region, extent, control, modify, groupby, do_flip, ignore_group_order,
                   dup_by_region, E, ooe, weight_name, coverage_norm, ignore_diags, W, store_stripes, max_windows,
                   trans=False, rescale_to=None, local=False):
    """``pileup_region`` = ``pos_stream`` -> ``_stream_snips`` -> ``accumulate_stream`` (coolpup.py:1285-1358).

    ``trans``: ``region`` / ``extent`` are pairs (region1, region2); the matrix is the rectangular region1 x region2
    block, ``E`` a scalar (1126-1128) and the diagonal mask is not applied (1141)."""
    if trans:
        (lo_rel, hi_rel), (lo_rel2, hi_rel2) = extent
        region, region2 = region
        nb2 = hi_rel2 - lo_rel2
    else:
        lo_rel, hi_rel = extent
        lo_rel2, region2 = lo_rel, region
        nb2 = hi_rel - lo_rel
    nb = hi_rel - lo_rel
    out = {"ROI": {}, "control": {}}
    log = {"st1": [], "st2": [], "kind": [], "flip": [], "group": []}
    bigdata = None
    empty = np.zeros((W, W))
    n_seen = 0
    stop = False
    frames = trans_frames(cc, region, region2, control, modify) if trans else cc.region_frames(region, control, modify)
    for fr in frames:
        cols = {c: fr[c].tolist() for c in fr.columns}  # to_dict(records) gives native python scalars
        nrow = len(fr)
        for i in range(nrow):
            if max_windows is not None and n_seen >= max_windows:
                stop = True
                break
            n_seen += 1
            rec = {c: cols[c][i] for c in cols}
            group = tuple(rec[c] for c in groupby) if groupby else "all"  # assign_groups (54-75)
            flip = bool(rec.get("flip", False)) if do_flip else False
            log["st1"].append(rec["stBin1"])
            log["st2"].append(rec["stBin2"])
            log["kind"].append(0 if rec["kind"] == "ROI" else 1)
            log["flip"].append(int(flip))
            log["group"].append(key_repr(group))
            if bigdata is None:  # get_data (1024-1057) + NaN-weight masks (1081-1098)
                bigdata = clr.matrix(sparse=True, balance=weight_name).fetch(tuple(region), tuple(region2)).tocsr()
                if weight_name:
                    isnan = np.isnan(clr.bins()[weight_name].fetch(tuple(region)).values)
                    isnan2 = np.isnan(clr.bins()[weight_name].fetch(tuple(region2)).values)
                else:
                    isnan = np.zeros(nb, dtype=bool)
                    isnan2 = np.zeros(nb2, dtype=bool)
                cov = clr.bins()[coverage_norm].fetch(tuple(region)).values if coverage_norm else None
                cov2 = clr.bins()[coverage_norm].fetch(tuple(region2)).values if coverage_norm else None
            # region-relative bins and bounds test (1105-1114)
            s1, e1 = rec["stBin1"] - lo_rel, rec["endBin1"] - lo_rel
            s2, e2 = rec["stBin2"] - lo_rel2, rec["endBin2"] - lo_rel2
            if s1 < 0 or e1 > nb or s2 < 0 or e2 > nb2:
                continue
            data = bigdata[s1:e1, s2:e2].toarray().astype(float)  # 1115-1121
            data[isnan[s1:e1], :] = np.nan  # 1122-1123
            data[:, isnan2[s2:e2]] = np.nan
            ii = np.arange(s1, e1)[:, None]
            jj = np.arange(s2, e2)[None, :]
            exp_data = None
            if E is not None and trans:  # np.full(data.shape, exp_value) (1126-1128)
                exp_data = np.full(data.shape, E)
            elif E is not None:  # expected_selections[region][s1:e1, s2:e2] = E[|i-j|] (1130-1133)
                exp_data = E[np.abs(jj - ii)]
            if not trans:
                data[(jj - ii) < ignore_diags] = np.nan  # signed diagonal mask (1141-1149)
            snip = {"kind": rec["kind"], "group": group, "coordinates": [], "horizontal_stripe": [],
                    "vertical_stripe": [], "cov_start": np.zeros(W), "cov_end": np.zeros(W)}
            if coverage_norm:  # 1151-1153
                snip["cov_start"] = cov[s1:e1]
                snip["cov_end"] = cov2[s2:e2]
            if E is not None and ooe:  # 1154-1156
                with np.errstate(divide="ignore", invalid="ignore"):
                    data = data / exp_data
            snip["data"] = data
            exp_snip = None
            if E is not None and not ooe:  # bare expected block as a control snippet (1135-1139)
                exp_snip = dict(snip)
                exp_snip["kind"] = "control"
                exp_snip["data"] = exp_data
                exp_snip["coordinates"] = []
            if rescale_to is not None:  # 1159-1162
                snip = rescale_snip(snip, rescale_to, local, coverage_norm)
                if exp_snip is not None:
                    exp_snip = rescale_snip(exp_snip, rescale_to, local, coverage_norm)
                data = snip["data"]
            if store_stripes:  # 1164-1182
                cntr = int(np.floor(data.shape[0] / 2))
                snip["horizontal_stripe"] = np.array(data[cntr, :], dtype=float)
                snip["vertical_stripe"] = np.array(data[:, cntr][::-1], dtype=float)
                snip["coordinates"] = ".".join(str(rec[c]) for c in ("chrom1", "start1", "end1", "chrom2", "start2", "end2"))
            emitted = [snip]
            if exp_snip is not None:  # yielded right after its snippet (1190-1191)
                if not store_stripes:
                    exp_snip["horizontal_stripe"], exp_snip["vertical_stripe"] = [], []
                emitted.append(exp_snip)
            for s in emitted:
                if do_flip and flip:  # flip_snip_func (128-147): anti-transpose, optional group swap
                    s["data"] = np.rot90(np.flipud(s["data"]))
                    if ignore_group_order:
                        sw = dict(rec)
                        for c in list(rec):
                            if c.endswith("1") and c[:-1] + "2" in rec:
                                sw[c], sw[c[:-1] + "2"] = rec[c[:-1] + "2"], rec[c]
                        if groupby:
                            s["group"] = tuple(sw[c] for c in groupby)
                if dup_by_region:  # group_by_region (lib/puputils.py:218-223)
                    targets = [(rec["chrom1"], rec["start1"], rec["end1"]), (rec["chrom2"], rec["start2"], rec["end2"])]
                else:
                    targets = [s["group"]]
                for g in targets:
                    add_snip(out[s["kind"]], g if isinstance(g, str) else tuple(g), s)
        if stop:
            break
    # "all" (1272-1282)
    empty_pup = {"data": empty, "horizontal_stripe": [], "vertical_stripe": [], "n": 0, "num": empty,
                 "cov_start": np.zeros(W), "cov_end": np.zeros(W), "coordinates": []}
    # NB the reference's sum_pups rebinds ``pup["data"] = np.nan_to_num(pup["data"])`` on BOTH of its inputs
    # (lib/puputils.py:97-98), so building "all" leaves every group of the region with NaN -> 0 and
    # +inf -> 1.797e308 in its own ``data`` (verified against the real reference: by-window groups that live in one
    # region only still show 1.797e308 / num)
    def all_of(groups):
        tot = reduce(sum_pups, groups.values(), empty_pup)
        for p in groups.values():
            p["data"] = np.nan_to_num(p["data"])
        return tot

    if "all" not in out["ROI"]:
        out["ROI"]["all"] = all_of(out["ROI"])
    if control or (E is not None and not ooe):
        if "all" not in out["control"]:
            out["control"]["all"] = all_of(out["control"])
    for k in ("st1", "st2", "kind", "flip"):
        log[k] = np.asarray(log[k], dtype=np.int64)
    return out, log


# --------------------------------------------------------------------------- boundary-level oracle (C ABI)
def oracle_accumulate(nb, indptr, col, count, weight, expected, coverage, r0, c0, slot, W, ignore_diags, n_slots,
                      ooe=False, expctrl=False):
    """What ``pup_accumulate`` + ``pup_acc_export`` must return, computed the reference's way.

    Literal per-window restatement of ``_stream_snips`` (coolpup.py:1104-1157) and ``_add_snip``
    (lib/puputils.py:12-38) over explicit window arrays: dense ``W x W`` slice of the symmetric CSR,
    ``(w_row * w_col) * count`` balancing as cooler does it, NaN rows / columns for NaN weights, signed diagonal
    mask, optional divide by ``expected[|col - row|]``, then ``nansum`` / ``isfinite`` accumulation per slot.
    With ``expctrl`` the bare expected block of every window is accumulated too (coolpup.py:1135-1139).
    """
    indptr = np.asarray(indptr)
    mat = sparse.csr_matrix((np.asarray(count, dtype=np.float64), np.asarray(col), indptr), shape=(nb, nb))
    if weight is not None:
        w = np.asarray(weight, dtype=np.float64)
        coo = mat.tocoo()
        coo.data = w[coo.row] * w[coo.col] * coo.data  # cooler: bias1[row] * bias2[col] * data
        mat = coo.tocsr()
        isnan = np.isnan(w)
    else:
        isnan = np.zeros(nb, dtype=bool)
    out = {
        "sum": np.zeros((n_slots, W, W)), "num": np.zeros((n_slots, W, W), dtype=np.int64),
        "n": np.zeros(n_slots, dtype=np.int64), "cov_start": np.zeros((n_slots, W)), "cov_end": np.zeros((n_slots, W)),
        "exp_sum": np.zeros((n_slots, W, W)), "exp_num": np.zeros((n_slots, W, W), dtype=np.int64),
    }
    for i in range(len(r0)):
        s1, s2, s = int(r0[i]), int(c0[i]), int(slot[i])
        if s1 < 0 or s1 + W > nb or s2 < 0 or s2 + W > nb:
            continue
        data = mat[s1 : s1 + W, s2 : s2 + W].toarray().astype(float)
        data[isnan[s1 : s1 + W], :] = np.nan
        data[:, isnan[s2 : s2 + W]] = np.nan
        ii = np.arange(s1, s1 + W)[:, None]
        jj = np.arange(s2, s2 + W)[None, :]
        data[(jj - ii) < ignore_diags] = np.nan
        if expected is not None and (ooe or expctrl):
            e = np.asarray(expected)[np.abs(jj - ii)]
            if ooe:
                with np.errstate(divide="ignore", invalid="ignore"):
                    data = data / e
            else:
                out["exp_sum"][s] = np.nansum([out["exp_sum"][s], e], axis=0)
                out["exp_num"][s] += np.isfinite(e)
        out["sum"][s] = np.nansum([out["sum"][s], data], axis=0)
        out["num"][s] += np.isfinite(data)
        out["n"][s] += 1
        if coverage is not None:
            cv = np.asarray(coverage, dtype=float)
            out["cov_start"][s] = np.nansum([out["cov_start"][s], cv[s1 : s1 + W]], axis=0)
            out["cov_end"][s] = np.nansum([out["cov_end"][s], cv[s2 : s2 + W]], axis=0)
    return out


def oracle_accumulate_rescaled(nb, indptr, col, count, weight, expected, coverage, r0, c0, h, w, slot, mode, rescale_size,
                               ignore_diags, n_slots, ooe=False, local=False):
    """What ``pup_accumulate_rescaled`` + ``pup_acc_export`` must return: ``_stream_snips`` (coolpup.py:1104-1157) over
    windows of their own sizes ``h[i] x w[i]``, every snippet through ``_rescale_snip`` (1193-1234), then ``_add_snip``.
    ``mode[i] == 1``: the snippet is the bare expected block (the control snippets of expected with ooe=False)."""
    indptr = np.asarray(indptr)
    mat = sparse.csr_matrix((np.asarray(count, dtype=np.float64), np.asarray(col), indptr), shape=(nb, nb))
    if weight is not None:
        wt = np.asarray(weight, dtype=np.float64)
        coo = mat.tocoo()
        coo.data = wt[coo.row] * wt[coo.col] * coo.data
        mat = coo.tocsr()
        isnan = np.isnan(wt)
    else:
        isnan = np.zeros(nb, dtype=bool)
    rs = int(rescale_size)
    out = {"sum": np.zeros((n_slots, rs, rs)), "num": np.zeros((n_slots, rs, rs), dtype=np.int64),
           "n": np.zeros(n_slots, dtype=np.int64), "cov_start": np.zeros((n_slots, rs)), "cov_end": np.zeros((n_slots, rs))}
    for i in range(len(r0)):
        s1, s2, hh, ww, s = int(r0[i]), int(c0[i]), int(h[i]), int(w[i]), int(slot[i])
        if s1 < 0 or s1 + hh > nb or s2 < 0 or s2 + ww > nb:
            continue
        ii = np.arange(s1, s1 + hh)[:, None]
        jj = np.arange(s2, s2 + ww)[None, :]
        if mode is not None and int(mode[i]) == 1:
            data = np.asarray(expected, dtype=float)[np.abs(jj - ii)]
        else:
            data = mat[s1 : s1 + hh, s2 : s2 + ww].toarray().astype(float)
            data[isnan[s1 : s1 + hh], :] = np.nan
            data[:, isnan[s2 : s2 + ww]] = np.nan
            data[(jj - ii) < ignore_diags] = np.nan
            if ooe:
                with np.errstate(divide="ignore", invalid="ignore"):
                    data = data / np.asarray(expected)[np.abs(jj - ii)]
        snip = {"data": data}
        if coverage is not None:
            cv = np.asarray(coverage, dtype=float)
            snip["cov_start"], snip["cov_end"] = cv[s1 : s1 + hh], cv[s2 : s2 + ww]
        snip = rescale_snip(snip, rs, local, coverage is not None and hh > 0 and ww > 0)
        out["sum"][s] = np.nansum([out["sum"][s], snip["data"]], axis=0)
        out["num"][s] += np.isfinite(snip["data"])
        out["n"][s] += 1
        if coverage is not None and hh > 0 and ww > 0:
            out["cov_start"][s] = np.nansum([out["cov_start"][s], snip["cov_start"]], axis=0)
            out["cov_end"][s] = np.nansum([out["cov_end"][s], snip["cov_end"]], axis=0)
    return out

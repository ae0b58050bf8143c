"""coolpuppy-compatible Python API on top of the B200 pile-up kernel.

Same public names, arguments and output DataFrame as the reference module
``coolpuppy/coolpup.py`` (``CoordCreator`` 150-749, ``PileUpper`` 752-1919,
``pileup`` 1922-2279), so ``plotpup`` / the CLI / ``save_pileup_df`` can consume
the result unchanged.  What differs is how the work is done:

* window coordinates are generated as numpy arrays per view region
  (:mod:`coolpuppy_b200._coords`) instead of one Python dict per window;
* snippet extraction, balancing, the signed diagonal mask, the expected
  divide and the running sum / count all happen in one CUDA kernel per region
  (``libpileup_b200.so``, C ABI in ``include/pileup_b200.h``);
* regions are spread over GPUs (one process per GPU, see
  :mod:`coolpuppy_b200.multigpu`) instead of over ``multiprocessing`` workers,
  and the per-GPU accumulators are merged with a single all-reduce.

There is no CPU fallback.  Not supported (raises ``NotImplementedError``):
``trans``, ``rescale``, arbitrary ``postprocess_func`` / ``extra_sum_funcs``
callbacks ("out of scope").
"""
from __future__ import annotations

import logging
import os
import warnings
from functools import partial

import numpy as np
import pandas as pd

from . import _native
from ._coords import RegionWindows, build_region_windows, default_band_edges, natsorted
from .coolio import is_cooler

logger = logging.getLogger("coolpuppy")

__all__ = [
    "CoordCreator", "PileUpper", "pileup", "bin_distance_intervals", "assign_groups", "expand", "expand2D",
    "flip_mark_intervals_func", "group_by_region", "make_cooler_view", "make_viewframe",
]


# ------------------------------------------------------------------------------------------ small frame helpers
def make_cooler_view(clr):
    """One region per chromosome (cooltools.lib.common.make_cooler_view; used at coolpup.py:858, 2123)."""
    names = list(clr.chromnames)
    return pd.DataFrame({"chrom": names, "start": 0, "end": [int(clr.chromsizes[c]) for c in names], "name": names})


def make_viewframe(view_df, check_bounds=None):
    """bioframe.make_viewframe for DataFrame input (coolpup.py:860): 3 columns -> name = chrom."""
    view = view_df.copy()
    if "chrom" not in view.columns:
        view.columns = ["chrom", "start", "end", "name"][: view.shape[1]]
    if "name" not in view.columns:
        view["name"] = view["chrom"].values
    view = view[["chrom", "start", "end", "name"]].reset_index(drop=True)
    view["chrom"] = view["chrom"].astype(str)
    view["name"] = view["name"].astype(str)
    if view["name"].duplicated().any():
        raise ValueError("view names must be unique")
    if check_bounds is not None:
        for c, s, e in zip(view["chrom"], view["start"], view["end"]):
            if c not in check_bounds.index or s < 0 or e > int(check_bounds[c]) or s > e:
                raise ValueError(f"view region {c}:{s}-{e} is out of the cooler's bounds")
    return view


def bin_distance_intervals(intervals, band_edges="default"):
    """Annotate 2-D intervals with their ``distance_band`` (coolpup.py:28-51)."""
    if isinstance(band_edges, str) and band_edges == "default":
        band_edges = default_band_edges()
    band_edges = np.asarray(band_edges)
    ids = np.searchsorted(band_edges, intervals["distance"], side="right")
    intervals["distance_band"] = [tuple(band_edges[i - 1 : i + 1]) for i in ids]
    return intervals


def assign_groups(intervals, groupby=[]):
    """``group`` column from ``groupby`` columns (coolpup.py:54-75)."""
    if not groupby:
        intervals["group"] = "all"
    else:
        intervals["group"] = list(intervals[groupby].values)
    return intervals


def expand(intervals, flank, resolution, rescale_flank=None):
    """Window [exp_start, exp_end) around each feature centre (coolpup.py:78-91)."""
    if rescale_flank is not None:
        raise NotImplementedError("rescaled pile-ups are not supported by the B200 path")
    intervals = intervals.copy()
    c = np.floor(intervals["center"] / resolution)
    intervals["exp_start"] = c * resolution - flank
    intervals["exp_end"] = (c + 1) * resolution + flank
    return intervals


def expand2D(intervals, flank, resolution, rescale_flank=None):
    """Two-sided version of :func:`expand` (coolpup.py:94-115)."""
    if rescale_flank is not None:
        raise NotImplementedError("rescaled pile-ups are not supported by the B200 path")
    for side in ("1", "2"):
        c = np.floor(intervals["center" + side] / resolution)
        intervals["exp_start" + side] = c * resolution - flank
        intervals["exp_end" + side] = (c + 1) * resolution + flank
    return intervals


def flip_mark_intervals_func(intervals, flipby, flip_negative_strand, extra_func=None):
    """``flip`` column for 2-D intervals (coolpup.py:118-125)."""
    if flip_negative_strand:
        intervals["flip"] = np.where(intervals["strand1"] == "-", True, False)
    else:
        intervals["flip"] = intervals[f"{flipby}1"] > intervals[f"{flipby}2"]
    if extra_func is not None:
        intervals = extra_func(intervals)
    return intervals


def group_by_region(snip):
    """Marker for by-window grouping (lib/puputils.py:218-223).

    The reference duplicates every snippet into the groups of its two
    anchors with this per-snippet callback; here it is recognised by identity
    and performed on the window arrays instead.
    """
    raise NotImplementedError("group_by_region is handled natively; pass it as postprocess_func")


# ------------------------------------------------------------------------------------------ CoordCreator
class CoordCreator:
    """Window-coordinate generator with the reference's constructor (coolpup.py:150-257)."""

    def __init__(self, features, resolution, *, features_format="auto", flank=100000, rescale_flank=None,
                 chroms="all", minshift=10**5, maxshift=10**6, nshifts=10, mindist="auto", maxdist=None,
                 local=False, subset=0, trans=False, seed=None):
        self.intervals = features.copy()
        self.resolution = resolution
        self.features_format = features_format
        self.flank = flank
        self.rescale_flank = rescale_flank
        self.chroms = chroms
        self.minshift = minshift
        self.maxshift = maxshift
        self.nshifts = nshifts
        self.trans = trans
        if trans:
            raise NotImplementedError("trans pile-ups are not supported by the B200 path")
        if rescale_flank is not None:
            raise NotImplementedError("rescaled pile-ups are not supported by the B200 path")
        self.mindist = 2 * self.flank + 2 * self.resolution if mindist == "auto" else mindist
        self.maxdist = np.inf if maxdist is None else maxdist
        self.local = local
        self.subset = subset
        self.seed = seed
        self.process()

    def process(self):
        """Centre, filter, sort and bin the features (coolpup.py:259-385)."""
        df = self.intervals
        if self.features_format is None or self.features_format == "auto":
            if all(c in df.columns for c in ["chrom1", "start1", "end1", "chrom2", "start2", "end2"]):
                self.kind = "bedpe"
            elif all(c in df.columns for c in ["chrom", "start", "end"]):
                self.kind = "bed"
            else:
                raise ValueError(
                    "Can't determine kind of input, please specify and/or name columns correctly:"
                    "'chrom1', 'start1', 'end1', 'chrom2', 'start2', 'end2' for bedpe kind"
                    "'chrom', 'start', 'end' for bed kind"
                )
        else:
            self.kind = self.features_format
        if self.subset > 0:
            df = self._subset(df)
        res, flank = self.resolution, self.flank
        if self.kind == "bed":
            assert all(c in df.columns for c in ["chrom", "start", "end"]), "Column names must include chrom, start, and end"
            df["chrom"] = df["chrom"].astype(str)
            df["center"] = (df["start"] + df["end"]) / 2
            df = expand(df, flank, res)
        elif self.kind == "bedpe":
            assert all(
                c in df.columns for c in ["chrom1", "start1", "end1", "chrom2", "start2", "end2"]
            ), "Column names must include chrom1, start1, end1, chrom2, start2, and end2"
            df[["chrom1", "chrom2"]] = df[["chrom1", "chrom2"]].astype(str)
            df["center1"] = (df["start1"] + df["end1"]) / 2
            df["center2"] = (df["start2"] + df["end2"]) / 2
            df["distance"] = df["center2"] - df["center1"]
            ad = df["distance"].abs()
            df = df[(self.mindist <= ad) & (ad <= self.maxdist)].reset_index(drop=True)
            df = expand2D(df, flank, res)
        else:
            raise ValueError('kind can only be "bed" or "bedpe"')
        self.intervals = df
        if df.shape[0] == 0:
            warnings.warn("No regions in features (maybe all below mindist?), returning empty output", stacklevel=2)
            self.pos_stream = self.empty_stream
            self.final_chroms = []
            return
        if self.kind == "bedpe" and self.nshifts > 0:
            df["kind"] = "ROI"
        if self.kind == "bed":
            basechroms = set(df["chrom"])
        else:
            if self.local:
                raise ValueError("Can't make local with both sides of loops defined")
            basechroms = set(df["chrom1"]).intersection(set(df["chrom2"]))
        self.basechroms = natsorted(list(basechroms))
        if isinstance(self.chroms, str) and self.chroms == "all":
            self.final_chroms = natsorted(list(basechroms))
        else:
            self.final_chroms = natsorted(list(set(self.chroms).intersection(set(self.basechroms))))
        if len(self.final_chroms) == 0:
            raise ValueError(
                "No chromosomes are in common between the coordinate file and the cooler file. "
                'Are they in the same format, e.g. starting with "chr"?'
            )
        self.intervals = self._binnify(df)
        self.pos_stream = self.get_combinations if self.kind == "bed" else self.get_intervals_stream

    def _subset(self, df):
        if self.seed is not None:
            np.random.seed(self.seed)
        if 0 < self.subset < len(df):
            return df.sample(self.subset)
        return df

    def _binnify(self, df):
        """Sort and convert bp windows to bins (coolpup.py:489-527)."""
        res = self.resolution
        if self.kind == "bed":
            df = df.sort_values(["chrom", "start"])
            df["stBin"] = np.floor(df["exp_start"] / res).astype(int)
            df["endBin"] = np.ceil(df["exp_end"] / res).astype(int)
            df[["exp_start", "exp_end"]] = df[["stBin", "endBin"]].values * res
        else:
            df = df.sort_values(["chrom1", "chrom2", "start1", "start2"])
            for s in ("1", "2"):
                df["stBin" + s] = np.floor(df["exp_start" + s] / res).astype(int)
                df["endBin" + s] = np.ceil(df["exp_end" + s] / res).astype(int)
                df[["exp_start" + s, "exp_end" + s]] = df[["stBin" + s, "endBin" + s]].values * res
        return df

    # -- array form (what the GPU path consumes) --------------------------------------------------
    def region_windows(self, region, control=False) -> RegionWindows:
        """All windows of view region ``(chrom, start, end)`` as arrays, in the reference's emission order."""
        return build_region_windows(self, region, control)

    # -- reference-compatible generators (slow; for callers that iterate the stream themselves) ----
    def _stream(self, region_filter, control, groupby, modify_2Dintervals_func):
        raise NotImplementedError(
            "pos_stream generators are replaced by CoordCreator.region_windows(region, control) in the B200 path"
        )

    def get_combinations(self, *a, **k):
        return self._stream(*a, **k)

    def get_intervals_stream(self, *a, **k):
        return self._stream(*a, **k)

    def empty_stream(self, *args, **kwargs):
        yield from ()


# ------------------------------------------------------------------------------------------ group bookkeeping
class _GroupTable:
    """Dense integer ids for group keys, shared by all regions (and ranks) of one run."""

    def __init__(self, cc, groupby, by_window):
        self.cc = cc
        self.groupby = list(groupby)
        self.by_window = by_window
        self._uniques = {}  # column -> list of values; code = position

    def _codes(self, col, values):
        """Global codes of ``values`` for group column ``col`` (codes are stable across regions)."""
        if col not in self._uniques:
            self._uniques[col] = self._global_uniques(col)
        uni = self._uniques[col]
        if uni is None:  # unknown globally (callback-made column): grow on demand
            uni = self._uniques[col] = {"_dynamic": True, "vals": [], "index": {}}
        if isinstance(uni, dict):
            codes, u = pd.factorize(pd.Series(list(values), dtype=object))
            m = np.empty(len(u), dtype=np.int64)
            for j, v in enumerate(u):
                if v not in uni["index"]:
                    uni["index"][v] = len(uni["vals"])
                    uni["vals"].append(v)
                m[j] = uni["index"][v]
            return m[codes]
        codes, u = pd.factorize(pd.Series(values))
        lut = pd.Series(np.arange(len(uni), dtype=np.int64), index=pd.Index(uni))
        m = lut.reindex(u).values
        if np.isnan(m.astype(float)).any():
            raise ValueError(f"group column {col!r}: value not present in the feature table")
        return m.astype(np.int64)[codes]

    def _global_uniques(self, col):
        df = self.cc.intervals
        if self.cc.kind == "bed" and col[-1] in "12" and col[:-1] in df.columns:
            base = col[:-1]
        elif col in df.columns:
            base = col
        else:
            return None
        vals = pd.unique(df[base])
        try:
            return sorted(vals.tolist())
        except TypeError:
            return vals.tolist()

    def value(self, col, code):
        uni = self._uniques[col]
        if isinstance(uni, dict):
            return uni["vals"][code]
        return uni[code]


def _band_ids(distance, edges):
    return np.searchsorted(edges, distance, side="right").astype(np.int64)


# ------------------------------------------------------------------------------------------ PileUpper
class PileUpper:
    """Pile-up engine with the reference's constructor (coolpup.py:752-997), running on a B200.

    Extra keyword (not in the reference): ``device`` -- CUDA device ordinal
    (default: ``LOCAL_RANK`` or 0).  ``nproc`` is accepted and recorded but
    unused: parallelism comes from the GPU(s).
    """

    def __init__(self, clr, CC, *, view_df=None, clr_weight_name="weight", expected=False,
                 expected_value_col="balanced.avg", ooe=True, control=False, coverage_norm=False, rescale=False,
                 rescale_size=99, flip_negative_strand=False, ignore_diags=2, store_stripes=False, nproc=1,
                 device=None):
        if not is_cooler(clr):
            raise TypeError("clr must be a cooler.Cooler or coolpuppy_b200.coolio.Cooler/MemCooler")
        self.clr = clr
        self.resolution = self.clr.binsize
        self.CC = CC
        assert self.resolution == self.CC.resolution
        for k in ("intervals", "features_format", "flank", "rescale_flank", "chroms", "minshift", "maxshift", "nshifts",
                  "trans", "mindist", "maxdist", "local", "subset", "seed", "kind", "basechroms", "final_chroms"):
            if hasattr(CC, k):
                setattr(self, k, getattr(CC, k))
        self.clr_weight_name = clr_weight_name
        self.expected = expected
        self.expected_value_col = expected_value_col
        self.ooe = ooe
        self.control = control
        self.pad_bins = self.CC.flank // self.resolution
        self.coverage_norm = coverage_norm
        self.rescale = rescale
        self.rescale_size = rescale_size
        self.flip_negative_strand = flip_negative_strand
        self.ignore_diags = ignore_diags
        self.store_stripes = store_stripes
        self.nproc = nproc
        self.ignore_group_order = False
        self._device = int(os.environ.get("LOCAL_RANK", "0")) if device is None else int(device)
        if rescale:
            raise NotImplementedError("rescaled pile-ups are not supported by the B200 path")
        if self.CC.flank % self.resolution != 0:
            raise ValueError("flank must be a multiple of the cooler's bin size")  # reference fails on shape mismatch

        if view_df is None:
            self.view_df = make_cooler_view(clr)
        else:
            self.view_df = make_viewframe(view_df, check_bounds=clr.chromsizes)
        self._expected_values = {}
        if self.expected is not None and self.expected is not False:
            exp = self.expected
            exp = exp[exp["region1"].isin(self.view_df["name"]) & exp["region2"].isin(self.view_df["name"])]
            if self.control:
                warnings.warn("Can't do both expected and control shifts; defaulting to expected", stacklevel=2)
                self.control = False
            exp = exp[exp["region1"] == exp["region2"]].reset_index(drop=True)
            for c in ("region1", "region2", "dist", self.expected_value_col):
                if c not in exp.columns:
                    raise ValueError("provided expected is not valid")
            for name in self.view_df["name"]:
                vals = exp.loc[(exp["region1"] == name), self.expected_value_col].values.astype(np.float64)
                self._expected_values[name] = vals  # E[d] in table row order (ExpectedSnipper.select, 907-916)
            self.expected_df = exp
            self.expected = True
        self.view_df = self.view_df.set_index("name")
        self.view_df_extents = {}
        for region_name, region in self.view_df.iterrows():
            lo, hi = self.clr.extent((region["chrom"], region["start"], region["end"]))
            chroffset = self.clr.offset(region["chrom"])
            self.view_df_extents[region_name] = lo - chroffset, hi - chroffset
        self.chroms = natsorted(list(set(self.CC.final_chroms) & set(self.clr.chromnames)))
        self.view_df = self.view_df[self.view_df["chrom"].isin(self.chroms)]
        if self.view_df["chrom"].unique().shape[0] == 0:
            raise ValueError(
                "No chromosomes are in common between the coordinate file and the cooler file. "
                'Are they in the same format, e.g. starting with "chr"?'
            )
        if self.coverage_norm is True:
            self.coverage_norm = "cov_tot_raw"
        elif self.coverage_norm == "cis":
            self.coverage_norm = "cov_cis_raw"
        elif self.coverage_norm == "total":
            self.coverage_norm = "cov_tot_raw"
        if self.coverage_norm and self.coverage_norm not in self.clr.bins().columns:
            if self.coverage_norm in ("cov_cis_raw", "cov_tot_raw"):
                raise NotImplementedError(
                    f"{self.coverage_norm} is not stored in the cooler; computing and storing coverage "
                    "(cooltools.coverage) is outside the B200 path -- run `cooltools coverage --store` first"
                )
            raise ValueError(f"coverage_norm {self.coverage_norm} not found in cooler bins")
        if self.coverage_norm and self.clr_weight_name:
            raise ValueError("Can't do coverage normalization when clr_weight_name is provided")
        self.empty_outmap = self.make_outmap()

    # -- small reference-compatible helpers -------------------------------------------------------
    def make_outmap(self):
        return np.zeros((2 * self.pad_bins + 1, 2 * self.pad_bins + 1))

    def get_data(self, region1, region2=None):
        """Region matrix as scipy CSR (coolpup.py:1024-1057); the GPU path uses :meth:`_region_arrays` instead."""
        r1 = self.view_df.loc[region1]
        r2 = r1 if region2 is None else self.view_df.loc[region2]
        return self.clr.matrix(sparse=True, balance=self.clr_weight_name).fetch(
            (r1["chrom"], r1["start"], r1["end"]), (r2["chrom"], r2["start"], r2["end"])
        ).tocsr()

    def _region_arrays(self, region_name):
        """Host arrays of one view region: symmetric CSR of raw counts + per-bin vectors."""
        r = self.view_df.loc[region_name]
        lo, hi = self.clr.extent((r["chrom"], r["start"], r["end"]))
        nb = hi - lo
        upper = False
        if hasattr(self.clr, "region_upper_csr"):
            indptr, col, cnt = self.clr.region_upper_csr(lo, hi)  # indexed as stored (pup_region_create_upper)
            upper = True
        else:  # a real cooler.Cooler
            m = self.clr.matrix(sparse=True, balance=False).fetch((r["chrom"], r["start"], r["end"])).tocsr()
            m.sort_indices()
            indptr, col, cnt = m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data.astype(np.int32)
        weight = cov = exp = None
        region = (r["chrom"], r["start"], r["end"])
        if self.clr_weight_name:
            weight = np.ascontiguousarray(self.clr.bins()[self.clr_weight_name].fetch(region).values, dtype=np.float64)
        if self.coverage_norm:
            cov = np.ascontiguousarray(self.clr.bins()[self.coverage_norm].fetch(region).values, dtype=np.float64)
        if self.expected is True:
            e = self._expected_values[region_name]
            exp = np.full(nb, np.nan)
            exp[: min(nb, len(e))] = e[:nb]
        return nb, indptr, col, cnt, weight, exp, cov, upper

    # -- the hot path -----------------------------------------------------------------------------
    def _plan(self, groupby, ignore_group_order, modify_2Dintervals_func, postprocess_func):
        """Resolve flip / grouping options exactly like pileupsWithControl (coolpup.py:1431-1493)."""
        by_window = postprocess_func is group_by_region
        if postprocess_func is not None and not by_window:
            raise NotImplementedError("arbitrary postprocess_func callbacks cannot run inside the CUDA kernel")
        flipby = None
        flip = False
        if self.flip_negative_strand:
            flipby = "strand"
            flip = True
            if ignore_group_order:
                if self.local:
                    raise ValueError("ignore_group_order doesn't make sense for local pileups")
                elif self.kind == "bedpe":
                    raise ValueError("ignore_group_order doesn't make sense for bedpe files")
                elif groupby:
                    warnings.warn("flip_negative_strand and ignore_group_order leads to combining strands, not other groups")
        elif ignore_group_order and groupby:
            if self.local:
                raise ValueError("ignore_group_order doesn't make sense for local pileups")
            if self.kind == "bedpe":
                raise ValueError("ignore_group_order doesn't make sense for bedpe files")
            groups = np.array(groupby)
            filt = [f"{g}1" in groups and f"{g}2" in groups for g in [x[:-1] for x in groups]]
            groups_filtered = np.sort(groups[filt])
            if ignore_group_order is True:
                fb = list(set(g[:-1] for g in groups_filtered))
            elif isinstance(ignore_group_order, str):
                fb = [ignore_group_order]
            elif len(ignore_group_order) == 1:
                fb = list(ignore_group_order)
            else:
                fb = list(set(g[:-1] for g in ignore_group_order))
            if len(fb) == 1 and f"{fb[0]}1" in groups_filtered:
                flipby = fb[0]
            else:
                raise ValueError("Ambiguous ignore_group_order, please provide str or list of two strings which are in groupby")
            flip = True
        elif ignore_group_order and not groupby:
            warnings.warn("Need to specify groupby for ignore_group_order")
        return dict(by_window=by_window, flip=flip, flipby=flipby, groupby=list(groupby),
                    ignore_group_order=ignore_group_order, modify=modify_2Dintervals_func)

    def _region_group_codes(self, rw: RegionWindows, plan, table: _GroupTable):
        """(flip flags, list of (target code arrays)) of a region's windows; codes are tuples of ints per column."""
        n = len(rw)
        modify = plan["modify"]
        band_edges = None
        if isinstance(modify, partial) and modify.func is bin_distance_intervals:
            band_edges = modify.keywords.get("band_edges", "default")
            if isinstance(band_edges, str) and band_edges == "default":
                band_edges = default_band_edges()
            band_edges = np.asarray(band_edges)
        elif modify is not None:
            # user callback: materialise the reference's DataFrame, let the callback annotate it
            fr = modify(rw.to_frame())
            if len(fr) != n:
                raise NotImplementedError("modify_2Dintervals_func must not add or drop rows in the B200 path")
            rw.frame = fr.reset_index(drop=True)
        flipf = np.zeros(n, dtype=bool)
        if plan["flip"] and n:
            if self.flip_negative_strand:
                flipf = np.asarray(rw.column("strand1") == "-")
            else:
                fb = plan["flipby"]
                flipf = np.asarray(rw.column(fb + "1") > rw.column(fb + "2"))
        swap = flipf if (plan["flip"] and plan["ignore_group_order"]) else None
        cols = []
        if plan["by_window"]:
            return flipf, None, band_edges
        for g in plan["groupby"]:
            if g == "distance_band" and band_edges is not None:
                cols.append(("band", _band_ids(rw.distance, band_edges)))
            else:
                cols.append((g, table._codes(g, rw.column(g, swap=swap)) if n else np.zeros(0, dtype=np.int64)))
        return flipf, cols, band_edges

    def _prepare(self, groupby, ignore_group_order, modify_2Dintervals_func, postprocess_func, regions=None,
                 dist=None):
        """Host phase: window arrays, group dictionary and accumulator slots of my view regions (no GPU needed)."""
        plan = self._plan(groupby, ignore_group_order, modify_2Dintervals_func, postprocess_func)
        W = 2 * self.pad_bins + 1
        table = _GroupTable(self.CC, plan["groupby"], plan["by_window"])
        region_names = list(self.view_df.index) if regions is None else list(regions)
        my_regions = region_names if dist is None else dist.my_items(region_names, self._region_cost)
        do_control = bool(self.control)
        expctrl = bool(self.expected is True and not self.ooe)
        built = []
        for ri, name in enumerate(region_names):
            r = self.view_df.loc[name]
            if name not in my_regions:
                if do_control and self.CC.nshifts > 0:
                    # another rank's region: still draw its control shifts so that every rank consumes the
                    # np.random stream exactly like the reference's serial (nproc=1) run
                    self.CC.region_windows((r["chrom"], r["start"], r["end"]), control=True)
                continue
            rw = self.CC.region_windows((r["chrom"], r["start"], r["end"]), control=do_control)
            if len(rw) == 0:
                continue
            flipf, cols, band_edges = self._region_group_codes(rw, plan, table)
            lo_rel, hi_rel = self.view_df_extents[name]
            nb = hi_rel - lo_rel
            r0 = rw.st1 - lo_rel
            c0 = rw.st2 - lo_rel
            valid = (r0 >= 0) & (r0 + W <= nb) & (c0 >= 0) & (c0 + W <= nb)
            if plan["by_window"]:
                # every window goes to the groups of both anchors (group_by_region, lib/puputils.py:218-223)
                ident = self._feature_ident(rw)
                keys = np.stack([ident[rw.idx1], ident[rw.idx2]], axis=1)[:, :, None]  # [n, 2, 1]
            elif cols:
                keys = np.stack([c for _, c in cols], axis=1)[:, None, :]  # [n, 1, ncols]
            else:
                keys = None
            built.append(dict(index=ri, name=name, rw=rw, r0=r0, c0=c0, valid=valid, flip=flipf, keys=keys,
                              band_edges=band_edges, colnames=[c for c, _ in cols] if cols else []))
        # group dictionary: unique keys in order of first appearance (region order, stream order)
        groups, gids, all_pos = self._assign_group_ids(built, plan, table, dist)
        nk = 2 if do_control else 1
        nf = 2 if plan["flip"] else 1
        for b, gid in zip(built, gids):
            rw = b["rw"]
            if gid.ndim == 2:  # by-window: two targets per window
                b["w_r0"] = np.repeat(b["r0"], 2)
                b["w_c0"] = np.repeat(b["c0"], 2)
                kind = np.repeat(rw.kind, 2)
                flip = np.repeat(b["flip"], 2)
                g = gid.reshape(-1)
                b["targets"] = 2
            else:
                b["w_r0"], b["w_c0"], kind, flip, g = b["r0"], b["c0"], rw.kind, b["flip"], gid
                b["targets"] = 1
            b["gid"] = g
            b["slot"] = (g * nk + kind.astype(np.int64)) * nf + flip.astype(np.int64)
        flags = 0
        if self.expected is True and self.ooe:
            flags |= _native.PUP_F_OOE
        if expctrl:
            flags |= _native.PUP_F_EXPCTRL
        if self.coverage_norm:
            flags |= _native.PUP_F_COVERAGE
        return dict(plan=plan, W=W, built=built, groups=groups, all_pos=all_pos, nk=nk, nf=nf,
                    n_slots=max(1, len(groups)) * nk * nf, flags=flags, do_control=do_control, expctrl=expctrl)

    def _run(self, groupby, ignore_group_order, modify_2Dintervals_func, postprocess_func, regions=None,
             dist=None):
        """Accumulate all (or the given) view regions on the GPU; returns the merged ROI / control pile-ups."""
        _native.require_device()
        job = self._prepare(groupby, ignore_group_order, modify_2Dintervals_func, postprocess_func, regions, dist)
        W, n_slots, flags = job["W"], job["n_slots"], job["flags"]
        stride = _native.acc_stride(W)
        acc = _native.alloc_accumulator(n_slots * stride, self._device)
        stream = _native.current_stream(self._device)
        self._last_stats = {"windows": 0, "launches": 0, "regions": 0}
        for b in job["built"]:
            nb, indptr, col, cnt, weight, exp, cov, upper = self._region_arrays(b["name"])
            region = _native.Region(self._device, nb, indptr, col, cnt, weight, exp, cov,
                                    ignore_diags=self.ignore_diags, flags=flags, stream=stream, upper=upper)
            self._last_stats["launches"] += int(_native.lib().pup_last_launches())
            try:
                nv = region.accumulate(
                    np.ascontiguousarray(b["w_r0"], dtype=np.int32), np.ascontiguousarray(b["w_c0"], dtype=np.int32),
                    np.ascontiguousarray(b["slot"], dtype=np.int32), W, n_slots, flags, acc,
                    stream=stream, want_n_valid=True)
                if self.store_stripes:
                    # per-ROI centre row / column (coolpup.py:1164-1182); a by-window pair is computed once
                    sel = np.nonzero(b["valid"] & (b["rw"].kind == 0))[0]
                    hor, ver = region.stripes(np.ascontiguousarray(b["r0"][sel], dtype=np.int32),
                                              np.ascontiguousarray(b["c0"][sel], dtype=np.int32), W, stream=stream)
                    b["stripes"] = (sel, hor, ver)
            finally:
                region.close()
            self._last_stats["windows"] += int(nv)
            self._last_stats["launches"] += int(_native.lib().pup_last_launches())
            self._last_stats["regions"] += 1
            n_roi = int(np.count_nonzero(b["valid"] & (b["rw"].kind == 0))) * b["targets"]
            if n_roi > 0:
                logger.info(f"{(b['name'], b['name'])}: {n_roi}")
        if dist is not None:
            dist.all_reduce(acc)
        out = _native.acc_export(acc, W, n_slots, device=self._device, stream=stream, want_expected=job["expctrl"],
                                 want_cov=bool(self.coverage_norm))
        plan = job["plan"]
        roi, ctrl = self._slots_to_pups(out, job["groups"], job["nk"], job["nf"], W, job["expctrl"], job["do_control"],
                                        grouped=bool(plan["groupby"]) or plan["by_window"], all_pos=job["all_pos"])
        if self.store_stripes:
            if dist is not None and dist.world_size > 1:
                raise NotImplementedError("store_stripes is a per-ROI output and is not gathered across ranks")
            self._attach_stripes(job, roi)
        return roi, ctrl

    def _attach_stripes(self, job, roi):
        """Per-group lists of stripes / coordinates in the reference's order: regions in view order; within a region
        the groups in order of first appearance, each in stream order; "all" concatenates the region's groups
        (sum_pups list concatenation, lib/puputils.py:105-107; coolpup.py:1272-1275)."""
        grouped = bool(job["plan"]["groupby"]) or job["plan"]["by_window"]
        groups = job["groups"]
        lists = {k: {"horizontal_stripe": [], "vertical_stripe": [], "coordinates": []} for k in roi}
        for b in job["built"]:
            if "stripes" not in b:
                continue
            sel, hor, ver = b["stripes"]
            rw = b["rw"]
            s_ = rw.sel
            if rw.paired:
                cols = [s_["chrom"].to_numpy()[rw.idx1[sel]], s_["start"].to_numpy()[rw.idx1[sel]], s_["end"].to_numpy()[rw.idx1[sel]],
                        s_["chrom"].to_numpy()[rw.idx2[sel]], s_["start"].to_numpy()[rw.idx2[sel]], s_["end"].to_numpy()[rw.idx2[sel]]]
            else:
                cols = [s_[c].to_numpy()[rw.idx1[sel]] for c in ("chrom1", "start1", "end1", "chrom2", "start2", "end2")]
            coords = [".".join(str(x.item() if isinstance(x, np.generic) else x) for x in row) for row in zip(*cols)]
            gid = b["gid"].reshape(len(rw), -1)[sel]  # [n_sel, targets]
            per_group = {}
            order = []
            for j in range(len(sel)):
                for g in gid[j]:
                    key = groups[int(g)] if grouped else "all"
                    if key not in per_group:
                        per_group[key] = []
                        order.append(key)
                    per_group[key].append(j)
            for key in order:
                if key not in lists:
                    continue
                for j in per_group[key]:
                    for dst in ([key, "all"] if grouped else [key]):
                        lists[dst]["horizontal_stripe"].append(hor[j])
                        lists[dst]["vertical_stripe"].append(ver[j])
                        lists[dst]["coordinates"].append(coords[j])
        for k, p in roi.items():
            p.update(lists[k])

    def _region_cost(self, name):
        """Predicted relative cost of a region (for LPT sharding): number of feature pairs."""
        r = self.view_df.loc[name]
        df = self.CC.intervals
        if self.CC.kind == "bedpe":
            n = int(((df["chrom1"] == r["chrom"]) & (df["start1"] >= r["start"]) & (df["end1"] < r["end"])).sum())
            return n
        n = int(((df["chrom"] == r["chrom"]) & (df["start"] >= r["start"]) & (df["end"] < r["end"])).sum())
        return n if self.local else n * (n - 1) // 2

    def _feature_ident(self, rw):
        """Global integer identity of each feature of the region table for by-window grouping."""
        if not hasattr(self, "_ident_index"):
            df = self.CC.intervals
            triples = pd.MultiIndex.from_arrays([df["chrom"].values, df["start"].values, df["end"].values])
            uniq = triples.unique()
            self._ident_values = list(uniq)
            self._ident_index = pd.Series(np.arange(len(uniq), dtype=np.int64), index=uniq)
        s = rw.sel
        mi = pd.MultiIndex.from_arrays([s["chrom"].values, s["start"].values, s["end"].values])
        return self._ident_index.reindex(mi).values.astype(np.int64)

    @staticmethod
    def _unique_rows(flat):
        """``(unique rows, index of each one's first occurrence, inverse)`` of an integer key matrix, by hashing one
        mixed-radix int64 per row instead of ``np.unique(axis=0)``'s lexicographic sort."""
        n = flat.shape[0]
        if n == 0 or flat.shape[1] == 0:
            return flat[:0], np.zeros(0, dtype=np.int64), np.zeros(n, dtype=np.int64)
        lo = flat.min(axis=0)
        radix = (flat.max(axis=0) - lo + 1).astype(object)
        span = 1
        for r in radix:
            span *= int(r)
        if span >= 2**62:  # cannot happen with group codes; keep the exact (slow) path for safety
            uniq, idx, inv = np.unique(flat, axis=0, return_index=True, return_inverse=True)
            return uniq, idx, np.asarray(inv).reshape(-1)
        ck = np.zeros(n, dtype=np.int64)
        for j in range(flat.shape[1]):
            ck = ck * int(radix[j]) + (flat[:, j] - lo[j])
        inv, _ = pd.factorize(ck)
        first = np.empty(int(inv.max()) + 1, dtype=np.int64)
        first[inv[::-1]] = np.arange(n - 1, -1, -1, dtype=np.int64)  # the smallest index is written last
        return flat[first], first, inv.astype(np.int64)

    def _assign_group_ids(self, built, plan, table, dist):
        """Dense group ids.  Returns (groups, gids, all_pos): ``groups`` lists the group keys in the reference's
        row order (first valid ROI emission, regions in view order), ``all_pos`` is where the reference's
        ``"all"`` row sits among them (after the groups first seen in the first view region, coolpup.py:1272-1275,
        1511-1520)."""
        if not plan["groupby"] and not plan["by_window"]:
            return ["all"], [np.zeros(len(b["rw"]), dtype=np.int64) for b in built], 0
        first = {}  # code tuple -> (is_control_only, region index, position of first valid emission)
        for b in built:
            keys = b["keys"]
            ntarget = keys.shape[1]
            flat = keys.reshape(len(keys) * ntarget, -1)
            b["_flat"] = flat
            for ctrl_only, ok in ((0, b["valid"] & (b["rw"].kind == 0)), (1, b["valid"] & (b["rw"].kind != 0))):
                ok = np.repeat(ok, ntarget)
                if not ok.any():
                    continue
                uniq, idx, _ = self._unique_rows(flat[ok])
                pos = np.nonzero(ok)[0][idx]
                for u, p in zip(map(tuple, uniq.tolist()), pos.tolist()):
                    cand = (ctrl_only, b["index"], p)
                    if u not in first or cand < first[u]:
                        first[u] = cand
        if dist is not None:
            first = dist.merge_min(first)
        order = sorted(first, key=lambda u: first[u])
        lookup = {u: i for i, u in enumerate(order)}
        all_pos = sum(1 for u in order if first[u][0] == 0 and first[u][1] == 0)
        colnames = next((b["colnames"] for b in built), [])
        edges = next((b["band_edges"] for b in built if b["band_edges"] is not None), None)
        groups = []
        for u in order:
            if plan["by_window"]:
                c, s, e = self._ident_values[u[0]]
                groups.append((c, int(s), int(e)))
            else:
                vals = []
                for colname, code in zip(colnames, u):
                    if colname == "band":
                        vals.append(tuple(edges[code - 1 : code + 1]))
                    else:
                        vals.append(table.value(colname, code))
                groups.append(tuple(vals))
        gids = []
        for b in built:
            flat = b.pop("_flat")
            ntarget = b["keys"].shape[1]
            uniq, _, inv = self._unique_rows(flat)
            # keys that never occur in a valid window are skipped by the kernel anyway: park them in group 0
            m = np.array([lookup.get(tuple(u), 0) for u in uniq.tolist()], dtype=np.int64)
            gid = m[np.asarray(inv).reshape(-1)]
            gids.append(gid.reshape(-1, ntarget) if ntarget == 2 else gid)
        return groups, gids, all_pos

    def _slots_to_pups(self, out, groups, nk, nf, W, expctrl, do_control, grouped, all_pos=0):
        """Per-group pile-ups from per-slot accumulators; flipped slots are anti-transposed (coolpup.py:130)."""

        def antit(a):
            return a[..., ::-1, ::-1].swapaxes(-1, -2)

        def gather(field, g, kind):
            s0 = (g * nk + kind) * nf
            v = out[field][s0]
            if nf == 2:
                w = out[field][s0 + 1]
                v = v + (antit(w) if w.ndim == 2 else w)
            return v

        roi, ctrl = {}, {}
        for g, key in enumerate(groups):
            p = {"data": gather("sum", g, 0), "num": gather("num", g, 0), "n": int(gather("n", g, 0)),
                 "horizontal_stripe": [], "vertical_stripe": [], "coordinates": []}
            if "cov_start" in out:
                p["cov_start"] = gather("cov_start", g, 0)
                p["cov_end"] = gather("cov_end", g, 0)
            else:
                p["cov_start"] = np.zeros(W)
                p["cov_end"] = np.zeros(W)
            roi[key] = p
            if do_control:
                c = {"data": gather("sum", g, 1), "num": gather("num", g, 1), "n": int(gather("n", g, 1)),
                     "horizontal_stripe": [], "vertical_stripe": [], "coordinates": []}
                if "cov_start" in out:
                    c["cov_start"] = gather("cov_start", g, 1)
                    c["cov_end"] = gather("cov_end", g, 1)
                else:
                    c["cov_start"] = np.zeros(W)
                    c["cov_end"] = np.zeros(W)
                ctrl[key] = c
            elif expctrl:
                # bare expected blocks are Toeplitz, hence invariant under the anti-transpose flip
                s0 = g * nk * nf
                es = sum(out["exp_sum"][s0 + f] for f in range(nf))
                en = sum(out["exp_num"][s0 + f] for f in range(nf))
                ctrl[key] = {"data": es, "num": en, "n": p["n"], "cov_start": p["cov_start"].copy(),
                             "cov_end": p["cov_end"].copy(), "horizontal_stripe": [], "vertical_stripe": [],
                             "coordinates": []}
        # groups without any accumulated window do not exist in the reference's dictionaries
        n_first_region = sum(1 for k in list(roi)[:all_pos] if roi[k]["n"] > 0)
        for d in (roi, ctrl):
            for k in [k for k, p in d.items() if p["n"] == 0 and not (isinstance(k, str) and k == "all")]:
                del d[k]
        if grouped:  # "all" = sum over groups (coolpup.py:1272-1282), placed where the reference's DataFrame has it
            for which, d, present in (("roi", roi, True), ("ctrl", ctrl, do_control or expctrl)):
                if not present:
                    continue
                tot = {"data": np.zeros((W, W)), "num": np.zeros((W, W), dtype=np.int64), "n": 0,
                       "cov_start": np.zeros(W), "cov_end": np.zeros(W), "horizontal_stripe": [],
                       "vertical_stripe": [], "coordinates": []}
                for p in d.values():
                    tot["data"] = tot["data"] + np.nan_to_num(p["data"])
                    tot["num"] = tot["num"] + p["num"]
                    tot["n"] += p["n"]
                    tot["cov_start"] = tot["cov_start"] + p["cov_start"]
                    tot["cov_end"] = tot["cov_end"] + p["cov_end"]
                items = list(d.items())
                items.insert(min(n_first_region, len(items)), ("all", tot))
                d.clear()
                d.update(items)
        return roi, ctrl

    def pileup_region(self, region1, region2=None, groupby=[], modify_2Dintervals_func=None, postprocess_func=None,
                      extra_sum_funcs=None):
        """Accumulated pile-ups of one view region: ``{"ROI": {group: pup}, "control": {...}}`` (coolpup.py:1285-1358)."""
        if region2 is not None and region2 != region1:
            raise NotImplementedError("trans pile-ups are not supported by the B200 path")
        if extra_sum_funcs:
            raise NotImplementedError("extra_sum_funcs callbacks are not supported by the B200 path")
        roi, ctrl = self._run(groupby, self.ignore_group_order, modify_2Dintervals_func, postprocess_func,
                              regions=[region1])
        return {"ROI": roi, "control": ctrl}

    def pileupsWithControl(self, nproc=None, groupby=[], ignore_group_order=False, modify_2Dintervals_func=None,
                           postprocess_func=None, extra_sum_funcs=None, dist=None):
        """Pile-ups over all view regions with the reference's normalisation (coolpup.py:1360-1654).

        ``dist`` (not in the reference): a :class:`coolpuppy_b200.multigpu.RegionSharder`; every rank must call
        this method, regions are split over ranks and all ranks return the same DataFrame.
        """
        self.ignore_group_order = ignore_group_order
        if extra_sum_funcs:
            raise NotImplementedError("extra_sum_funcs callbacks are not supported by the B200 path")
        if len(self.chroms) == 0:
            return self.make_outmap(), 0
        roi, ctrl = self._run(groupby, ignore_group_order, modify_2Dintervals_func, postprocess_func, dist=dist)
        has_ctrl = bool(self.control or (self.expected is True and not self.ooe))
        if self.coverage_norm:
            roi = {k: _norm_coverage(p) for k, p in roi.items()}
            if self.control:
                ctrl = {k: _norm_coverage(p) for k, p in ctrl.items()}
            elif self.expected is True:
                warnings.warn("Expected can not be normalized to coverage", stacklevel=2)
        rows = {"group": [], "data": [], "control_n": [], "control_num": [], "n": [], "num": []}
        with np.errstate(divide="ignore", invalid="ignore"):
            for k, p in roi.items():
                data = p["data"] / p["num"]
                if has_ctrl:
                    c = ctrl.get(k)
                    if c is not None:
                        data = data / (c["data"] / c["num"])
                        rows["control_n"].append(c["n"])
                        rows["control_num"].append(c["num"])
                    else:
                        data = data * np.nan
                        rows["control_n"].append(np.nan)
                        rows["control_num"].append(np.nan)
                data = np.where(data == np.inf, np.nan, data)
                if self.local:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore", category=RuntimeWarning)
                        data = np.nanmean(np.dstack((data, data.T)), 2)
                rows["group"].append(k)
                rows["data"].append(data)
                rows["n"].append(p["n"])
                rows["num"].append(p["num"])
        if not has_ctrl:
            del rows["control_n"], rows["control_num"]
        if self.store_stripes:  # coolpup.py:1556-1600
            W = 2 * self.pad_bins + 1
            cntr = W // 2
            rows["coordinates"], rows["horizontal_stripe"], rows["vertical_stripe"] = [], [], []
            with np.errstate(divide="ignore", invalid="ignore"):
                if has_ctrl:
                    call = ctrl["all"]
                    cnorm = call["data"] / call["num"]
                    ch, cv = cnorm[cntr, :], cnorm[:, cntr][::-1]
                for k, p in roi.items():
                    hs = np.vstack(p["horizontal_stripe"]) if p["horizontal_stripe"] else np.zeros((0, W))
                    vs = np.vstack(p["vertical_stripe"]) if p["vertical_stripe"] else np.zeros((0, W))
                    if has_ctrl:
                        hs, vs = hs / ch, vs / cv
                    if self.local:  # numutils._copy_array_halves
                        vs[:, : cntr + 1] = np.fliplr(vs[:, cntr:])
                        hs[:, : cntr + 1] = np.fliplr(hs[:, cntr:])
                    rows["coordinates"].append(np.vstack([c.split(".") for c in p["coordinates"]])
                                               if p["coordinates"] else np.zeros((0, 6), dtype=str))
                    rows["horizontal_stripe"].append(hs)
                    rows["vertical_stripe"].append(vs)
        n = roi["all"]["n"]
        normalized_roi = pd.DataFrame({k: _objcol(v) if k in ("group", "data", "num", "control_num") else v
                                       for k, v in rows.items()
                                       if k not in ("coordinates", "horizontal_stripe", "vertical_stripe")})
        if self.store_stripes:
            for c in ("coordinates", "horizontal_stripe", "vertical_stripe"):
                normalized_roi[c] = _objcol(rows[c])
        if groupby:
            glist = [("all",) * len(groupby) if (isinstance(i, str) and i == "all") else i
                     for i in normalized_roi["group"].to_list()]
            for j, val in enumerate(groupby):
                normalized_roi.insert(0, val, _objcol([g[j] for g in glist]))
        logger.info(f"Total number of piled up windows: {int(n)}")
        for name, attr in self._annotation_items():
            if isinstance(attr, list):
                attr = str(attr)
            normalized_roi[name] = attr
        return normalized_roi

    def _annotation_items(self):
        """Attribute columns in the reference's order (coolpup.py:1628-1653)."""
        names = ["clr", "resolution", "flank", "rescale_flank", "chroms", "minshift", "maxshift", "nshifts", "trans",
                 "mindist", "maxdist", "local", "subset", "seed", "clr_weight_name", "expected", "expected_value_col",
                 "ooe", "control", "pad_bins", "coverage_norm", "rescale", "rescale_size", "flip_negative_strand",
                 "ignore_diags", "store_stripes", "nproc", "ignore_group_order"]
        for nme in names:
            v = getattr(self, nme)
            if nme == "clr":
                v = os.path.abspath(self.clr.filename)
            yield nme, v

    def pileupsByStrandWithControl(self, nproc=None, groupby=[], ignore_group_order=False, dist=None):
        """By-strand wrapper (coolpup.py:1656-1694)."""
        normalized_pileups = self.pileupsWithControl(
            nproc=nproc, groupby=["strand1", "strand2"] + groupby, ignore_group_order=ignore_group_order, dist=dist)
        normalized_pileups.insert(
            0, "orientation",
            (normalized_pileups["strand1"].astype(str) + normalized_pileups["strand2"].astype(str)).replace({"allall": "all"}),
        )
        return normalized_pileups

    def pileupsByWindowWithControl(self, nproc=None, dist=None):
        """By-window wrapper (coolpup.py:1696-1755): one row per feature plus ``all``."""
        if self.local:
            raise ValueError("Cannot do by-window pileups for local")
        pups = self.pileupsWithControl(nproc=nproc, postprocess_func=group_by_region, dist=dist)
        trip = [("all", -1, -1) if (isinstance(g, str) and g == "all") else g for g in pups["group"]]
        coords = pd.DataFrame(trip, index=pups.index, columns=["chrom", "start", "end"])
        pups = pd.concat([coords, pups], axis=1).drop(columns="group")
        pups[["start", "end"]] = pups[["start", "end"]].astype(int)
        order = {c: i for i, c in enumerate(pd.unique(self.view_df["chrom"]))}
        rank = pups["chrom"].map(lambda c: order.get(c, len(order)))
        pups = pups.assign(_r=rank.values).sort_values(["_r", "start", "end"], kind="stable").drop(columns="_r")
        return pups.reset_index(drop=True)

    def _distance_edges(self, distance_edges):
        if not (isinstance(distance_edges, str) and distance_edges == "default"):
            if not all(isinstance(n, (int, np.integer)) for n in distance_edges):
                raise ValueError("Distance edges must be integers")
            distance_edges = list(np.sort(distance_edges))
            for _ in range(len(distance_edges)):
                if np.min(distance_edges) < self.mindist:
                    distance_edges[int(np.argmin(distance_edges))] = self.mindist
                else:
                    break
        return distance_edges

    @staticmethod
    def _separation(x):
        if isinstance(x, str) and x == "all":
            return x
        if len(x) == 2:
            return f"{x[0]/1000000}Mb-\n{x[1]/1000000}Mb"
        return f"{x[0]/1000000}Mb+"

    def pileupsByDistanceWithControl(self, nproc=None, distance_edges="default", groupby=[], ignore_group_order=False,
                                     dist=None):
        """By-distance wrapper (coolpup.py:1757-1833)."""
        if self.local:
            raise ValueError("Cannot do by-distance pileups for local")
        bin_func = partial(bin_distance_intervals, band_edges=self._distance_edges(distance_edges))
        pups = self.pileupsWithControl(nproc=nproc, modify_2Dintervals_func=bin_func,
                                       groupby=["distance_band"] + groupby, ignore_group_order=ignore_group_order,
                                       dist=dist)
        keep = [not (isinstance(x, tuple) and len(x) == 0) for x in pups["distance_band"]]
        pups = pups.loc[keep, :].reset_index(drop=True)
        pups.insert(0, "separation", pups["distance_band"].apply(self._separation))
        i = np.where(pups["separation"] == "all")[0]
        pups = pd.concat([_sort_rows(pups.drop(i), ["distance_band"]), pups.iloc[i, :]], ignore_index=True)
        return pups.reset_index(drop=True)

    def pileupsByStrandByDistanceWithControl(self, nproc=None, distance_edges="default", groupby=[],
                                             ignore_group_order=False, dist=None):
        """By-strand-by-distance wrapper (coolpup.py:1835-1919)."""
        bin_func = partial(bin_distance_intervals, band_edges=self._distance_edges(distance_edges))
        pups = self.pileupsWithControl(nproc=nproc, modify_2Dintervals_func=bin_func,
                                       groupby=["strand1", "strand2", "distance_band"] + groupby,
                                       ignore_group_order=ignore_group_order, dist=dist)
        pups.insert(0, "orientation",
                    (pups["strand1"].astype(str) + pups["strand2"].astype(str)).replace({"allall": "all"}))
        keep = [not (isinstance(x, tuple) and len(x) == 0) for x in pups["distance_band"]]
        pups = pups.loc[keep, :].reset_index(drop=True)
        pups.insert(0, "separation", pups["distance_band"].apply(self._separation))
        i = np.where(pups["separation"] == "all")[0]
        pups = pd.concat([_sort_rows(pups.drop(i), ["orientation", "distance_band"]), pups.iloc[i, :]],
                         ignore_index=True)
        return pups.reset_index(drop=True)


def _objcol(values):
    """A pandas column whose cells hold arbitrary objects (arrays, tuples)."""
    arr = np.empty(len(values), dtype=object)
    for i, v in enumerate(values):
        arr[i] = v
    return pd.Series(arr, dtype=object)


def _sort_rows(df, cols):
    """``df.sort_values(cols)`` where a column may hold tuples."""
    keys = [tuple(row) for row in df[cols].itertuples(index=False, name=None)]
    order = sorted(range(len(keys)), key=lambda i: keys[i])
    return df.iloc[order]


def _norm_coverage(pup):
    """Coverage normalisation of an accumulated pile-up (lib/puputils.py:168-190)."""
    cov = np.outer(pup["cov_start"], pup["cov_end"])
    with np.errstate(divide="ignore", invalid="ignore"):
        cov = cov / np.nanmean(cov)
        data = pup["data"] / cov
    data[np.isnan(data)] = 0
    out = dict(pup)
    out["data"] = data
    return out


# ------------------------------------------------------------------------------------------ pileup()
def pileup(clr, features, features_format="bed", view_df=None, expected_df=None, expected_value_col="balanced.avg",
           clr_weight_name="weight", flank=100000, minshift=10**5, maxshift=10**6, nshifts=0, ooe=True,
           mindist="auto", maxdist=None, min_diag=2, subset=0, by_window=False, by_strand=False, by_distance=False,
           groupby=[], ignore_group_order=False, flip_negative_strand=False, local=False, coverage_norm=False,
           trans=False, rescale=False, rescale_flank=1, rescale_size=99, store_stripes=False, nproc=1, seed=None,
           device=None, dist=None):
    """One-call pile-up with the reference's signature (coolpup.py:1922-2279).

    Extra keywords: ``device`` (CUDA ordinal) and ``dist`` (multi-GPU region sharder).
    """
    if by_distance is not False:
        if local:
            raise ValueError("Can't do local pileups by distance, please specify only one of those arguments")
        if isinstance(by_distance, np.ndarray):
            try:
                distance_edges = [int(i) for i in by_distance]
            except Exception as e:
                raise ValueError("Distance bin edges have to be an iterable of integers or convertable to integers") from e
            by_distance = True
        elif by_distance is True or (isinstance(by_distance, str) and by_distance == "default"):
            distance_edges = "default"
            by_distance = True
        else:
            raise ValueError("Invalid by_distance value, should be either True, 'default' or a list of integers")
    if not rescale:
        rescale_flank = None
    if seed is not None:
        np.random.seed(seed)
    if view_df is None:
        view_df = make_cooler_view(clr)
    else:
        try:
            make_viewframe(view_df, check_bounds=clr.chromsizes)
        except Exception as e:
            raise ValueError("view_df is not a valid viewframe or incompatible") from e
    control = nshifts > 0
    if expected_df is None:
        expected_value_col = None
    else:
        need = ["region1", "region2", "dist", expected_value_col]
        if not all(c in expected_df.columns for c in need):
            raise ValueError("provided expected is not valid")
    if mindist is None:
        mindist = "auto"
    if maxdist is None:
        maxdist = np.inf
    if rescale and rescale_size % 2 == 0:
        raise ValueError("Please provide an odd rescale_size")
    chroms = list(view_df["chrom"].unique())
    if by_window:
        if features_format != "bed":
            raise ValueError("Can't make by-window pileups without making combinations")
        if local:
            raise ValueError("Can't make local by-window pileups")

    CC = CoordCreator(features=features, resolution=clr.binsize, features_format=features_format, flank=flank,
                      rescale_flank=rescale_flank, chroms=chroms, minshift=minshift, maxshift=maxshift, nshifts=nshifts,
                      mindist=mindist, maxdist=maxdist, local=local, subset=subset, seed=seed, trans=trans)
    PU = PileUpper(clr=clr, CC=CC, view_df=view_df, clr_weight_name=clr_weight_name, expected=expected_df,
                   expected_value_col=expected_value_col, ooe=ooe, control=control, coverage_norm=coverage_norm,
                   rescale=rescale, rescale_size=rescale_size, flip_negative_strand=flip_negative_strand,
                   ignore_diags=min_diag, store_stripes=store_stripes, nproc=nproc, device=device)
    if by_window:
        pups = PU.pileupsByWindowWithControl(dist=dist)
        flags = (True, False, False)
        if groupby:
            warnings.warn("by-window not compatible with additional groupby")
    elif by_strand and by_distance:
        pups = PU.pileupsByStrandByDistanceWithControl(nproc=nproc, distance_edges=distance_edges, groupby=groupby,
                                                       ignore_group_order=ignore_group_order, dist=dist)
        flags = (False, True, True)
    elif by_strand:
        pups = PU.pileupsByStrandWithControl(groupby=groupby, ignore_group_order=ignore_group_order, dist=dist)
        flags = (False, True, False)
    elif by_distance:
        pups = PU.pileupsByDistanceWithControl(nproc=nproc, distance_edges=distance_edges, groupby=groupby,
                                               ignore_group_order=ignore_group_order, dist=dist)
        flags = (False, False, True)
    else:
        pups = PU.pileupsWithControl(groupby=groupby, ignore_group_order=ignore_group_order, dist=dist)
        flags = (False, False, False)
    pups["by_window"], pups["by_strand"], pups["by_distance"] = flags
    pups["groupby"] = [groupby] * pups.shape[0]
    pups["expected"] = [False if (e is None or e is False or (isinstance(e, float) and np.isnan(e))) else e for e in pups["expected"]]
    pups["cooler"] = os.path.splitext(os.path.basename(clr.filename))[0]
    return pups

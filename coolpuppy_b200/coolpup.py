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
``rescale``, arbitrary ``postprocess_func`` / ``extra_sum_funcs`` callbacks
("out of scope").
"""
from __future__ import annotations

import logging
import os
import warnings
from functools import partial

import numpy as np
import pandas as pd

from . import _native
from ._coords import RegionWindows, build_region_windows, build_trans_windows, default_band_edges, natsorted
from .coolio import is_cooler

logger = logging.getLogger("coolpuppy")
_LAST_STATS = {}  # statistics of the most recent PileUpper run in this process (bench.py reads them after pileup())

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


def _expand_scale(start, end, scale):
    """``bioframe.expand(df, scale=scale)`` (third-party; restated from its source): grow every interval about its
    midpoint to ``scale`` times its length, round half to even, keep the integer dtype."""
    start = np.asarray(start)
    end = np.asarray(end)
    pads = 0.5 * (scale - 1) * (end - start)
    return np.round(start - pads).astype(start.dtype), np.round(end + pads).astype(end.dtype)


def expand(intervals, flank, resolution, rescale_flank=None):
    """Window [exp_start, exp_end) around each feature centre, or the feature itself grown by ``rescale_flank`` of its
    size on either side for rescaled pile-ups (coolpup.py:78-91)."""
    intervals = intervals.copy()
    if rescale_flank is None:
        c = np.floor(intervals["center"] / resolution)
        intervals["exp_start"] = c * resolution - flank
        intervals["exp_end"] = (c + 1) * resolution + flank
    else:
        intervals["exp_start"], intervals["exp_end"] = _expand_scale(intervals["start"].values, intervals["end"].values,
                                                                     2 * rescale_flank + 1)
    return intervals


def expand2D(intervals, flank, resolution, rescale_flank=None):
    """Two-sided version of :func:`expand` (coolpup.py:94-115)."""
    for side in ("1", "2"):
        if rescale_flank is None:
            c = np.floor(intervals["center" + side] / resolution)
            intervals["exp_start" + side] = c * resolution - flank
            intervals["exp_end" + side] = (c + 1) * resolution + flank
        else:
            a, b = _expand_scale(intervals["start" + side].values, intervals["end" + side].values, 2 * rescale_flank + 1)
            intervals["exp_start" + side], intervals["exp_end" + side] = a, b
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
        if isinstance(mindist, str) and mindist == "auto":
            self.mindist = 2 * self.flank + 2 * self.resolution
        else:
            self.mindist = mindist
            if self.trans:  # coolpup.py:243-246
                warnings.warn("Ignoring mindist when using trans", stacklevel=2)
                self.mindist = 0
        if maxdist is None:
            self.maxdist = np.inf
        else:
            self.maxdist = maxdist
            if self.trans:  # coolpup.py:250-253
                warnings.warn("Ignoring maxdist when using trans", stacklevel=2)
                self.maxdist = np.inf
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
            df = expand(df, flank, res, self.rescale_flank)
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
            df = expand2D(df, flank, res, self.rescale_flank)
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
            basechroms = set(df["chrom"].unique().tolist())
        else:
            if self.local:
                raise ValueError("Can't make local with both sides of loops defined")
            if self.trans:  # coolpup.py:341-345
                basechroms = set(df["chrom1"].unique().tolist() + df["chrom2"].unique().tolist())
            else:
                basechroms = set(df["chrom1"].unique().tolist()).intersection(set(df["chrom2"].unique().tolist()))
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
        if self.trans and self.local:
            raise ValueError("Cannot do local with trans=True")
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

    def region_windows_trans(self, region1, region2, control=False) -> RegionWindows:
        """Windows between two view regions on different chromosomes (trans pile-ups)."""
        return build_trans_windows(self, region1, region2, control)

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
def _isnull(v):
    return v is None or (isinstance(v, float) and v != v)


class _GroupTable:
    """Dense integer codes for the values of group columns, identical on all regions and ranks of one run.

    A column that exists in the feature table gets its code dictionary from the whole table up front (sorted values,
    missing values last).  A column that only a user ``modify_2Dintervals_func`` creates is *dynamic*: codes are handed
    out in order of appearance while the regions are laid out and :meth:`finalize` replaces them by positions in the
    dictionary merged over all ranks.
    """

    def __init__(self, cc):
        self.cc = cc
        self._static = {}   # column -> (values list, has_na)
        self._dynamic = {}  # column -> {"vals": [...], "index": {...}, "lut": None}

    def _global_uniques(self, col):
        df = self.cc.intervals
        if self.cc.kind == "bed" and col[-1] in "12" and col[:-1] in df.columns:
            base = col[:-1]
        elif col in df.columns:
            base = col
        else:
            return None
        vals = pd.unique(df[base]).tolist()
        has_na = any(_isnull(v) for v in vals)
        vals = [v for v in vals if not _isnull(v)]
        try:
            vals = sorted(vals)
        except TypeError:
            pass
        return vals, has_na

    def is_dynamic(self, col):
        if col not in self._static and col not in self._dynamic:
            u = self._global_uniques(col)
            if u is None:
                self._dynamic[col] = {"vals": [], "index": {}, "lut": None}
            else:
                self._static[col] = u
        return col in self._dynamic

    def codes(self, col, values):
        """Codes of ``values`` for group column ``col``; a missing value (NaN / None) is a group of its own."""
        if self.is_dynamic(col):
            dyn = self._dynamic[col]

            def code_of(k, v):
                if k not in dyn["index"]:
                    dyn["index"][k] = len(dyn["vals"])
                    dyn["vals"].append(v)
                return dyn["index"][k]

            codes, u = pd.factorize(pd.Series(list(values), dtype=object), use_na_sentinel=True)
            m = np.full(len(u) + 1, -1, dtype=np.int64)
            for j, v in enumerate(u):
                m[j] = code_of(v, v)
            if (codes < 0).any():
                m[-1] = code_of("__na__", None)
            return m[codes]  # code -1 (missing value) picks the last entry
        vals, has_na = self._static[col]
        codes, u = pd.factorize(pd.Series(values), use_na_sentinel=True)
        lut = pd.Series(np.arange(len(vals), dtype=np.int64), index=pd.Index(vals, dtype=object) if len(vals) else None)
        m = lut.reindex(pd.Index(list(u), dtype=object)).values.astype(np.float64) if len(u) else np.zeros(0)
        if np.isnan(m).any():
            raise ValueError(f"group column {col!r}: value not present in the feature table")
        m = np.append(m.astype(np.int64), len(vals))  # missing values -> the extra last code
        if (codes < 0).any() and not has_na:
            raise ValueError(f"group column {col!r}: missing value not present in the feature table")
        return m[codes]

    def finalize(self, dist):
        """Make the dynamic dictionaries global: merge them over the ranks and remember local -> global codes."""
        for col, dyn in sorted(self._dynamic.items()):
            lists = [dyn["vals"]] if dist is None else dist.all_gather_object(dyn["vals"])
            merged, seen = [], set()
            for lst in lists:
                for v in lst:
                    k = "__na__" if _isnull(v) else v
                    if k not in seen:
                        seen.add(k)
                        merged.append(v)
            na = [v for v in merged if _isnull(v)]
            rest = [v for v in merged if not _isnull(v)]
            try:
                rest = sorted(rest)
            except TypeError:
                pass
            merged = rest + na[:1]
            pos = {("__na__" if _isnull(v) else v): i for i, v in enumerate(merged)}
            dyn["lut"] = np.array([pos["__na__" if _isnull(v) else v] for v in dyn["vals"]], dtype=np.int64)
            dyn["global"] = merged

    def remap(self, col, codes):
        if col in self._dynamic:
            lut = self._dynamic[col]["lut"]
            return lut[codes] if len(lut) else codes
        return codes

    def radix(self, col):
        if col in self._dynamic:
            return max(1, len(self._dynamic[col]["global"]))
        vals, has_na = self._static[col]
        return len(vals) + 1

    def value(self, col, code):
        if col in self._dynamic:
            return self._dynamic[col]["global"][code]
        vals, _ = self._static[col]
        return vals[code] if code < len(vals) else np.nan


def _band_ids(distance, edges):
    return np.searchsorted(edges, distance, side="right").astype(np.int64)


def _first_positions(keys, ok, pos):
    """{key: smallest pos} over the entries with ``ok`` (keys: int64 array)."""
    if not ok.any():
        return {}
    k, p = keys[ok], pos[ok]
    order = np.argsort(k, kind="stable")  # pos ascends inside the input, so the first of every run is the minimum
    ks, ps = k[order], p[order]
    head = np.ones(len(ks), dtype=bool)
    head[1:] = ks[1:] != ks[:-1]
    # positions are not necessarily ascending for strided parts of by-window targets: take the true minimum
    mins = np.minimum.reduceat(ps, np.nonzero(head)[0])
    return dict(zip(ks[head].tolist(), mins.tolist()))


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
        if rescale:  # coolpup.py:970-975
            if getattr(self, "rescale_flank", None) is None:
                raise ValueError("Cannot use rescale without setting rescale_flank")
            if rescale_size % 2 == 0:
                raise ValueError("Please provide an odd rescale_size")
            if store_stripes or self.trans:
                raise NotImplementedError("rescaled pile-ups with store_stripes or trans are not supported by the B200 path")
        elif self.CC.flank % self.resolution != 0:
            raise ValueError("flank must be a multiple of the cooler's bin size")  # reference fails on shape mismatch
        elif getattr(self, "rescale_flank", None) is not None:
            raise ValueError("rescale_flank without rescale=True gives windows of different sizes")

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
            if self.trans:  # one scalar per region pair (get_expected_trans, coolpup.py:999-1005)
                for c in ("region1", "region2", self.expected_value_col):
                    if c not in exp.columns:
                        raise ValueError("provided expected is not valid")
                self.expected_df = exp.reset_index(drop=True)
                self.expected = True
            else:
                exp = exp[exp["region1"] == exp["region2"]].reset_index(drop=True)
                for c in ("region1", "region2", "dist", self.expected_value_col):
                    if c not in exp.columns:
                        raise ValueError("provided expected is not valid")
                by_region = {k: v.values.astype(np.float64)
                             for k, v in exp.groupby("region1", sort=False)[self.expected_value_col]}
                for name in self.view_df["name"]:
                    # E[d] in table row order (ExpectedSnipper.select, 907-916)
                    self._expected_values[name] = by_region.get(name, np.zeros(0))
                self.expected_df = exp
                self.expected = True
        self.view_df = self.view_df.set_index("name")
        self.view_df_extents = {}
        for region_name, region in self.view_df.iterrows():
            lo, hi = self.clr.extent((region["chrom"], region["start"], region["end"]))
            chroffset = self.clr.offset(region["chrom"])
            self.view_df_extents[region_name] = lo - chroffset, hi - chroffset
        if self.expected is True and not self.trans:
            # cooltools.lib.checks.is_valid_expected(..., verify_cooler=clr) at coolpup.py:875-906: every view region
            # needs one expected row per diagonal; a missing region would silently give an all-NaN pile-up here
            for region_name, (lo_rel, hi_rel) in self.view_df_extents.items():
                have = len(self._expected_values.get(region_name, ()))
                if have == 0:
                    raise ValueError(f"provided expected is not valid: no cis rows for view region {region_name!r}")
                if have < hi_rel - lo_rel:
                    raise ValueError(f"provided expected is not valid: {have} diagonals for view region "
                                     f"{region_name!r} of {hi_rel - lo_rel} bins")
        self.chroms = natsorted(list(set(self.CC.final_chroms) & set(self.clr.chromnames)))
        self.view_df = self.view_df[self.view_df["chrom"].isin(self.chroms)]
        if self.view_df["chrom"].unique().shape[0] == 0:
            raise ValueError(
                "No chromosomes are in common between the coordinate file and the cooler file. "
                'Are they in the same format, e.g. starting with "chr"?'
            )
        if self.trans and self.view_df["chrom"].unique().shape[0] < 2:
            raise ValueError("Trying to do trans with fewer than two chromosomes")
        if self.coverage_norm is True:
            self.coverage_norm = "cov_tot_raw"
        elif self.coverage_norm == "cis":
            self.coverage_norm = "cov_cis_raw"
        elif self.coverage_norm == "total":
            self.coverage_norm = "cov_tot_raw"
        if (self.coverage_norm in ("cov_cis_raw", "cov_tot_raw") and self.coverage_norm not in self.clr.bins().columns
                and hasattr(self.clr, "add_bin_column")):
            # the reference computes the coverage with cooltools and STORES it in the cooler (coolpup.py:955-963);
            # here it is computed the same way (expected.coverage, pinned against cooltools-made columns) and kept in
            # memory
            from .expected import coverage as _coverage

            try:
                cis, tot = _coverage(self.clr, ignore_diags=self.ignore_diags)
                self.clr.add_bin_column("cov_cis_raw", cis)
                self.clr.add_bin_column("cov_tot_raw", tot)
            except NotImplementedError:
                pass
        if self.coverage_norm and self.coverage_norm not in self.clr.bins().columns:
            if self.coverage_norm in ("cov_cis_raw", "cov_tot_raw"):
                raise NotImplementedError(
                    f"{self.coverage_norm} is not stored in the cooler and this cooler object cannot take an in-memory "
                    "column -- run `cooltools coverage --store` first"
                )
            raise ValueError(f"coverage_norm {self.coverage_norm} not found in cooler bins")
        if self.coverage_norm and self.clr_weight_name:
            raise ValueError("Can't do coverage normalization when clr_weight_name is provided")
        self.empty_outmap = self.make_outmap()
        self._cost_cache = {}

    # -- small reference-compatible helpers -------------------------------------------------------
    def _out_size(self):
        """Side of the output pile-up: rescale_size, or 2 * pad + 1 bins (make_outmap, coolpup.py:1007-1022)."""
        return int(self.rescale_size) if self.rescale else 2 * self.pad_bins + 1

    def make_outmap(self):
        return np.zeros((self._out_size(), self._out_size()))

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

    def _band_edges(self, plan):
        modify = plan["modify"]
        if isinstance(modify, partial) and modify.func is bin_distance_intervals:
            e = modify.keywords.get("band_edges", "default")
            if isinstance(e, str) and e == "default":
                e = default_band_edges()
            return np.asarray(e)
        return None

    def _region_group_codes(self, rw: RegionWindows, plan, table: _GroupTable):
        """(flip flags, [(column, codes)]) of a region's windows; ``None`` instead of the list for by-window."""
        n = len(rw)
        modify = plan["modify"]
        band_edges = plan["band_edges"]
        if modify is not None and band_edges is None:
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
        if plan["by_window"]:
            return flipf, None
        cols = []
        for g in plan["groupby"]:
            if g == "distance_band" and band_edges is not None:
                cols.append((g, _band_ids(rw.distance, band_edges)))
            else:
                table.is_dynamic(g)  # registers the column even when this region has no windows
                cols.append((g, table.codes(g, rw.column(g, swap=swap)) if n else np.zeros(0, dtype=np.int64)))
        return flipf, cols

    def _region_nnz(self, name):
        """Stored pixels of a view region's rows (cheap, from the cooler's row index) or None."""
        off = getattr(self.clr, "_bin1_offset", None)
        if off is None:
            return None
        r = self.view_df.loc[name]
        lo, hi = self.clr.extent((r["chrom"], r["start"], r["end"]))
        return int(off[hi] - off[lo])

    # pile-up time model of one window (SM cycles on B200, fitted to configs[3]): the sparse kernel
    # pays per strip run and per stored pixel, the dense-band kernel reads every cell of the window
    _COST_SPARSE_PER_ROW, _COST_SPARSE_PER_PIXEL, _COST_DENSE_PER_CELL, _DENSE_FILL = 3.8, 0.52, 0.1465, 0.20
    _COST_PER_WINDOW = 95.0  # sort, chunk plan and count kernels
    _COST_PER_REGION = 3.5e7  # launch ramps and tails of one region's kernels (~0.12 ms of the whole GPU)

    @staticmethod
    def _fit_poisson_scale(nb, nnz):
        """lambda0 with sum_s (nb - s) * (1 - exp(-lambda0 / s)) = nnz: stored pixels of a matrix whose counts are
        Poisson(lambda0 / separation) -- the occupancy model of the sharding costs (bisection, s sampled)."""
        s_ = np.unique(np.round(np.geomspace(1, max(2, nb - 1), 400)).astype(np.int64))
        wgt = np.gradient(s_.astype(float)) * (nb - s_)
        lo_, hi_ = 1e-6, 1e9
        for _ in range(60):
            mid = np.sqrt(lo_ * hi_)
            if float(np.sum(wgt * -np.expm1(-mid / s_))) < nnz:
                lo_ = mid
            else:
                hi_ = mid
        return float(np.sqrt(lo_ * hi_))

    def _feature_costs(self, name):
        """Predicted pile-up time (relative units) of the windows anchored at every feature of a view region, in the
        order of the region's feature table: for bed pairs the pairs (k, l > k) count for their row anchor k.  The
        occupancy of a window at separation s is modelled as 1 - exp(-lambda0 / s) (contact density ~ 1 / separation,
        lambda0 fitted to the region's stored pixel count); a window costs a term per tile row plus a term per stored
        pixel in the sparse kernel, or a term per cell when it lies in the dense diagonal band (occupancy >= 20 %).
        The sum is the region's cost for the sharding; its running sum gives the cut points of window parts."""
        if name in self._cost_cache:
            return self._cost_cache[name]
        r = self.view_df.loc[name]
        df = self.CC.intervals
        W = self._out_size()
        lo, hi = self.view_df_extents[name]
        nb = max(2, hi - lo)
        nnz = self._region_nnz(name)
        lam = 1.0 if nnz is None else self._fit_poisson_scale(nb, max(1.0, float(nnz)))  # upper-triangle pixels
        reps = 1 + (self.CC.nshifts if self.control else 0)
        res = float(self.resolution)
        band_ok = not self.trans and self.ignore_diags >= 0  # the regions that get a dense band

        def cost_of(sep_bins):
            sep = np.maximum(np.abs(sep_bins), 1.0)
            fill = -np.expm1(-lam / sep)
            cost = self._COST_SPARSE_PER_ROW * W + self._COST_SPARSE_PER_PIXEL * W * W * fill
            if band_ok:
                cost = np.where((fill >= self._DENSE_FILL) & (sep >= W - 1), self._COST_DENSE_PER_CELL * W * W, cost)
            return (cost + self._COST_PER_WINDOW) * reps

        if self.CC.kind == "bedpe":
            m = ((df["chrom1"].values == r["chrom"]) & (df["chrom2"].values == r["chrom"])
                 & (df["start1"].values >= r["start"]) & (df["end1"].values < r["end"])
                 & (df["start2"].values >= r["start"]) & (df["end2"].values < r["end"]))
            out = cost_of(df["distance"].values[m] / res)
        else:
            m = (df["chrom"].values == r["chrom"]) & (df["start"].values >= r["start"]) & (df["end"].values < r["end"])
            c = df["center"].values[m] / res
            n = len(c)
            if self.local:
                out = cost_of(np.zeros(n))
            else:
                out = np.zeros(n)
                lo_d, hi_d = self.mindist / res, self.maxdist / res
                step = max(1, n // 2000)  # quadratic in the sites: evaluate every step-th anchor, interpolate
                ks = np.arange(0, n, step)
                for a in range(0, len(ks), 256):
                    kk = ks[a : a + 256]
                    d = np.abs(c[None, :] - c[kk, None])
                    keep = (np.arange(n)[None, :] > kk[:, None]) & (d >= lo_d) & (d <= hi_d)
                    out[kk] = np.where(keep, cost_of(d), 0.0).sum(axis=1)
                if step > 1:
                    out = np.interp(np.arange(n), ks, out[ks])
        out = np.asarray(out, dtype=np.float64)
        if len(out) and out.sum() > 0:
            out = out + self._COST_PER_REGION / len(out)
        self._cost_cache[name] = out
        return self._cost_cache[name]

    def _region_cost(self, name):
        return float(self._feature_costs(name).sum())

    def _pair_cost(self, name1, name2):
        """Relative cost of a trans region pair: the number of windows between the two regions."""
        df = self.CC.intervals

        def inside(name, side):
            r = self.view_df.loc[name]
            return ((df["chrom" + side].values == r["chrom"]) & (df["start" + side].values >= r["start"])
                    & (df["end" + side].values < r["end"]))

        if self.CC.kind == "bedpe":
            return float(np.count_nonzero(inside(name1, "1") & inside(name2, "2"))
                         + np.count_nonzero(inside(name1, "2") & inside(name2, "1"))) + 1e-3
        return float(np.count_nonzero(inside(name1, "")) * np.count_nonzero(inside(name2, ""))) + 1e-3

    def _part_ranges(self, name, my_parts, parts):
        """Feature-index ranges ``[(k_lo, k_hi)]`` (merged when adjacent) of window parts ``my_parts`` out of ``parts``:
        the region's features are cut at equal predicted cost.  A part holds the windows whose ROW anchor lies in its
        range (controls stay with their ROI pair), i.e. a band of matrix rows -- every rank streams only its band
        from HBM, unlike a strided split whose parts each read the whole matrix."""
        cost = self._feature_costs(name)
        n = len(cost)
        if parts <= 1 or n == 0:
            return [(0, n)]
        cum = np.concatenate([[0.0], np.cumsum(cost)])
        cuts = np.searchsorted(cum, cum[-1] * np.arange(1, parts) / parts, side="left")
        cuts = np.concatenate([[0], np.clip(cuts, 0, n), [n]])
        cuts = np.maximum.accumulate(cuts)
        ranges = []
        for p_ in sorted(my_parts):
            a, b = int(cuts[p_]), int(cuts[p_ + 1])
            if b <= a:
                continue
            if ranges and ranges[-1][1] == a:
                ranges[-1] = (ranges[-1][0], b)
            else:
                ranges.append((a, b))
        return ranges

    def _my_units(self, region_names, dist, splittable=True):
        """{region: [feature-index ranges]} of this rank + the predicted max / mean load over the ranks."""
        self._cost_cache = {}
        if dist is None or dist.world_size == 1:
            return {name: None for name in region_names}, 1.0
        if self.trans:
            units, imbalance = dist.my_units(region_names, [self._pair_cost(*n) for n in region_names], max_share=1e9)
            return {name: None for name, _, _ in units}, imbalance
        if splittable and hasattr(dist, "my_ranges"):
            # one contiguous, equal-cost piece of the (region, row anchor) sequence per rank
            return dist.my_ranges(region_names, [self._feature_costs(n) for n in region_names])
        costs = [self._region_cost(n) for n in region_names]
        units, imbalance = dist.my_units(region_names, costs, max_share=0.25 if splittable else 1e9)
        mine = {}
        for name, part, parts in units:
            mine.setdefault(name, (set(), parts))[0].add(part)
        return {name: (None if parts == 1 else self._part_ranges(name, ps, parts)) for name, (ps, parts) in mine.items()}, imbalance

    def _needs_exact_merge(self):
        """True when a pixel can become +inf (x / 0 with ooe): the reference's region / group merge then turns it
        into 1.797e308 (np.nan_to_num inside sum_pups, lib/puputils.py:97-98), which depends on how the partial
        pile-ups were grouped -- reproduced by :meth:`_exact_merge` from per-region accumulators."""
        if not (self.expected is True and self.ooe):
            return False
        if self.trans:
            return bool(np.any(self.expected_df[self.expected_value_col].values == 0))
        return any(np.any(v == 0) for v in self._expected_values.values())

    def _prepare(self, plan, regions=None, dist=None):
        """Host phase: window arrays, dense group keys and accumulator slots of my sharding units (no GPU needed)."""
        modify_2Dintervals_func = plan["modify"]
        W = self._out_size()
        table = _GroupTable(self.CC)
        region_names = self._unit_keys() if regions is None else list(regions)
        do_control = bool(self.control)
        expctrl = bool(self.expected is True and not self.ooe)
        splittable = not (self.trans or self.store_stripes
                          or (modify_2Dintervals_func is not None and plan["band_edges"] is None))
        my_units, imbalance = self._my_units(region_names, dist, splittable)
        built = []
        for ri, name in enumerate(region_names):
            if self.trans:
                ra, rb = self.view_df.loc[name[0]], self.view_df.loc[name[1]]
                region = ((ra["chrom"], ra["start"], ra["end"]), (rb["chrom"], rb["start"], rb["end"]))
                make = lambda draw_only=False: build_trans_windows(self.CC, region[0], region[1], do_control, draw_only=draw_only)
            else:
                r = self.view_df.loc[name]
                region = (r["chrom"], r["start"], r["end"])
                make = lambda draw_only=False: build_region_windows(self.CC, region, do_control, draw_only=draw_only)
            if name not in my_units:
                if do_control and self.CC.nshifts > 0:
                    # another rank's region: make its np.random draws (no window layout) so that every rank consumes
                    # the random stream exactly like the reference's serial (nproc=1) run
                    make(draw_only=True)
                continue
            rw = make()
            if len(rw) == 0:
                continue
            pos0 = np.arange(len(rw), dtype=np.int64)
            if my_units[name] is not None:  # my row bands of the region's windows (matrix replicated on the other owners)
                mask = np.zeros(len(rw), dtype=bool)
                for k_lo, k_hi in my_units[name]:
                    mask |= (rw.idx1 >= k_lo) & (rw.idx1 < k_hi)
                pos0 = pos0[mask]
                rw = rw.take(pos0)
                if len(rw) == 0:
                    continue
            flipf, cols = self._region_group_codes(rw, plan, table)
            if self.trans:
                # the rectangular region1 x region2 matrix is embedded in a square one of nb1 + nb2 bins (rows: region1,
                # columns: region2 behind an offset of nb1); a window must lie inside both regions (coolpup.py:1111-1114)
                (lo1, hi1), (lo2, hi2) = self.view_df_extents[name[0]], self.view_df_extents[name[1]]
                a, b = rw.st1 - lo1, rw.st2 - lo2
                valid = (a >= 0) & (a + W <= hi1 - lo1) & (b >= 0) & (b + W <= hi2 - lo2)
                r0 = np.where(valid, a, -1)  # the kernel drops windows with a negative corner
                c0 = b + (hi1 - lo1)
            else:
                lo_rel, hi_rel = self.view_df_extents[name]
                nb = hi_rel - lo_rel
                r0 = rw.st1 - lo_rel
                c0 = rw.st2 - lo_rel
                if self.rescale:  # windows of their features' own sizes (zoomed to rescale_size on the device)
                    hh, ww = rw.sizes()
                    valid = (r0 >= 0) & (r0 + hh <= nb) & (c0 >= 0) & (c0 + ww <= nb)
                else:
                    valid = (r0 >= 0) & (r0 + W <= nb) & (c0 >= 0) & (c0 + W <= nb)
            built.append(dict(index=ri, name=name, rw=rw, r0=r0, c0=c0, valid=valid, flip=flipf, cols=cols, pos0=pos0))
            if self.rescale:
                built[-1].update(h=hh, w=ww)
        # dense group keys: mixed radix over the group columns (identical on every rank), or the feature id (by-window)
        table.finalize(dist)
        if plan["by_window"]:
            self._feature_ident(None)
            colspec = [("window", len(self._ident_values))]
        else:
            colspec = []
            for g in plan["groupby"]:
                if g == "distance_band" and plan["band_edges"] is not None:
                    colspec.append((g, len(plan["band_edges"]) + 1))
                else:
                    table.is_dynamic(g)
                    colspec.append((g, table.radix(g)))
        n_keys = 1
        for _, radix in colspec:
            n_keys *= int(radix)
        if n_keys >= 2**31:
            raise ValueError("too many possible groups for dense accumulator slots")
        # kinds: ROI, and "control" when there are shifted controls -- or, for rescaled pile-ups with expected and
        # ooe=False, the bare expected blocks, which are zoomed like snippets and therefore need slots of their own
        nk = 2 if (do_control or (self.rescale and expctrl)) else 1
        nf = 2 if plan["flip"] else 1
        first = {}  # key -> (is_control_only, region index, position of the first valid emission)
        for b in built:
            rw = b["rw"]
            n = len(rw)
            if plan["by_window"]:
                # every window goes to the groups of both anchors (group_by_region, lib/puputils.py:218-223)
                ident1, ident2 = self._feature_ident(rw)
                key = np.stack([ident1[rw.idx1], ident2[rw.idx2]], axis=1).reshape(-1)  # [n * 2]
                ntarget = 2
            else:
                key = np.zeros(n, dtype=np.int64)
                for (g, radix), (_, codes) in zip(colspec, b["cols"]):
                    key = key * int(radix) + table.remap(g, codes)
                ntarget = 1
            b["targets"] = ntarget
            kind = np.repeat(rw.kind, ntarget).astype(np.int64)
            flip = np.repeat(b["flip"], ntarget).astype(np.int64)
            b["w_r0"] = np.repeat(b["r0"], ntarget)
            b["w_c0"] = np.repeat(b["c0"], ntarget)
            b["key"] = key
            b["slot"] = (key * nk + kind) * nf + flip
            if self.rescale:
                b["w_h"], b["w_w"] = np.repeat(b["h"], ntarget), np.repeat(b["w"], ntarget)
                b["mode"] = np.zeros(len(key), dtype=np.int64)
                if expctrl:  # every snippet is followed by its expected block as a control snippet (coolpup.py:1135-1139)
                    for f in ("w_r0", "w_c0", "w_h", "w_w"):
                        b[f] = np.concatenate([b[f], b[f]])
                    b["slot"] = np.concatenate([b["slot"], (key * nk + 1) * nf + flip])
                    b["mode"] = np.concatenate([b["mode"], np.ones(len(key), dtype=np.int64)])
                order = np.argsort(b["slot"], kind="stable")  # the rescale kernel flushes its tile when the slot changes
                for f in ("w_r0", "w_c0", "w_h", "w_w", "slot", "mode"):
                    b[f] = b[f][order]
            pos = np.repeat(b["pos0"], ntarget) * ntarget + np.tile(np.arange(ntarget), n)
            valid = np.repeat(b["valid"], ntarget)
            for ctrl_only, ok in ((0, valid & (kind == 0)), (1, valid & (kind != 0))):
                for k, p_ in _first_positions(key, ok, pos).items():
                    cand = (ctrl_only, b["index"], p_)
                    if k not in first or cand < first[k]:
                        first[k] = cand
        if dist is not None:
            first = dist.merge_min(first)
        flags = 0
        if self.expected is True and self.ooe:
            flags |= _native.PUP_F_OOE
        if expctrl:
            flags |= _native.PUP_F_EXPCTRL
        if self.coverage_norm:
            flags |= _native.PUP_F_COVERAGE
        if self.trans:
            flags |= _native.PUP_F_NODIAG  # no diagonal mask between chromosomes (coolpup.py:1141)
        if self.rescale and self.local:
            flags |= _native.PUP_F_LOCAL
        return dict(plan=plan, W=W, built=built, colspec=colspec, table=table, first=first, nk=nk, nf=nf, n_keys=n_keys,
                    n_slots=n_keys * nk * nf, flags=flags, do_control=do_control, expctrl=expctrl,
                    region_names=region_names, imbalance=imbalance)

    def _decode_key(self, job, key):
        """Group key tuple (as the reference builds it, coolpup.py:71-75) of dense key ``key``."""
        plan = job["plan"]
        if not plan["groupby"] and not plan["by_window"]:
            return "all"
        if plan["by_window"]:
            c, s_, e = self._ident_values[key]
            return (c, int(s_), int(e))
        vals = []
        for g, radix in reversed(job["colspec"]):
            key, code = divmod(key, int(radix))
            if g == "distance_band" and plan["band_edges"] is not None:
                vals.append(tuple(plan["band_edges"][code - 1 : code + 1]))
            else:
                vals.append(job["table"].value(g, code))
        return tuple(reversed(vals))

    def _unit_keys(self):
        """What pileup_region is mapped over (coolpup.py:1419-1429): the view regions, or for trans every pair of view
        regions on different chromosomes."""
        names = list(self.view_df.index)
        if not self.trans:
            return names
        import itertools

        return [(a, b) for a, b in itertools.combinations(names, 2)
                if self.view_df.loc[a, "chrom"] != self.view_df.loc[b, "chrom"]]

    def _expected_trans(self, region1, region2):
        """get_expected_trans (coolpup.py:999-1005): the one expected value of a region pair."""
        e = self.expected_df
        vals = e.loc[(e["region1"] == region1) & (e["region2"] == region2), self.expected_value_col]
        return float(vals.item())

    def _trans_arrays(self, name1, name2):
        """Host arrays of a trans region pair: the region1 x region2 block as the upper-right block of a square CSR of
        nb1 + nb2 bins, per-bin vectors concatenated, the scalar expected as a constant vector."""
        ra, rb = self.view_df.loc[name1], self.view_df.loc[name2]
        reg1, reg2 = (ra["chrom"], ra["start"], ra["end"]), (rb["chrom"], rb["start"], rb["end"])
        m = self.clr.matrix(sparse=True, balance=False).fetch(reg1, reg2).tocsr()
        m.sort_indices()
        nb1, nb2 = m.shape
        nb = nb1 + nb2
        indptr = np.concatenate([m.indptr, np.full(nb2, m.indptr[-1])]).astype(np.int32)
        col = (m.indices.astype(np.int64) + nb1).astype(np.int32)
        cnt = m.data
        if not np.issubdtype(cnt.dtype, np.integer):
            raise NotImplementedError("floating-point pixel counts are not supported by the CUDA path")
        cnt = np.ascontiguousarray(cnt, dtype=np.int32)

        def both(colname):
            f = self.clr.bins()[colname].fetch
            return np.ascontiguousarray(np.concatenate([f(reg1).values, f(reg2).values]), dtype=np.float64)

        weight = both(self.clr_weight_name) if self.clr_weight_name else None
        cov = both(self.coverage_norm) if self.coverage_norm else None
        exp = np.full(nb, self._expected_trans(name1, name2)) if self.expected is True else None
        return nb, indptr, col, cnt, weight, exp, cov, False

    def _region_kwargs(self, name, flags):
        if self.trans:
            nb, indptr, col, cnt, weight, exp, cov, upper = self._trans_arrays(*name)
            return dict(nb=nb, indptr=indptr, col=col, count=cnt, weight=weight, expected=exp, coverage=cov, ignore_diags=0,
                        flags=flags & (_native.PUP_F_OOE | _native.PUP_F_NODIAG), upper=False)
        nb, indptr, col, cnt, weight, exp, cov, upper = self._region_arrays(name)
        return dict(nb=nb, indptr=indptr, col=col, count=cnt, weight=weight, expected=exp, coverage=cov,
                    ignore_diags=self.ignore_diags, flags=flags & (_native.PUP_F_OOE | _native.PUP_F_NODIAG), upper=upper)

    def _device_windows_ok(self, plan):
        """Can the windows of this run be generated on the GPU (``pup_pair_windows_device``)?  Yes for bed features
        paired all-vs-all grouped by feature columns / strands / distance bands / by window; bedpe, local pile-ups,
        user callbacks and per-ROI stripes keep the host window builder."""
        if os.environ.get("PUP_DEVICE_WINDOWS", "1") == "0" or not _native.device_windows_supported():
            return False
        if self.CC.kind != "bed" or self.local or self.store_stripes or self.trans or self.rescale:
            return False
        if plan["modify"] is not None and plan["band_edges"] is None:
            return False
        df = self.CC.intervals
        for g in plan["groupby"]:
            if g == "distance_band" and plan["band_edges"] is not None:
                continue
            if not (g[-1] in "12" and g[:-1] in df.columns):
                return False
        if plan["flip"] and not self.flip_negative_strand and plan["flipby"] not in df.columns:
            return False
        if plan["flip"] and self.flip_negative_strand and "strand" not in df.columns:
            return False
        if self.control and self.CC.nshifts > 0:
            lo, hi = int(self.CC.minshift), int(self.CC.maxshift)
            if not (lo + 1 < hi and -2**31 < lo and hi < 2**31 and hi - 1 - lo < 2**32 - 1):
                return False
        return True

    def _prepare_device(self, plan, regions, dist):
        """Host phase of the device-window path: per view region only the feature table is touched (window bins,
        centres, per-feature group-key parts, kept pairs per offset); windows, control shifts, slots and the
        first-appearance order of the groups are produced on the GPU."""
        W = self._out_size()
        table = _GroupTable(self.CC)
        df = self.CC.intervals
        region_names = list(self.view_df.index) if regions is None else list(regions)
        do_control = bool(self.control) and self.CC.nshifts > 0
        expctrl = bool(self.expected is True and not self.ooe)
        my_units, imbalance = self._my_units(region_names, dist)
        # dense key space (static columns only on this path)
        if plan["by_window"]:
            self._feature_ident(None)
            colspec = [("window", len(self._ident_values))]
        else:
            colspec = []
            for g in plan["groupby"]:
                if g == "distance_band" and plan["band_edges"] is not None:
                    colspec.append((g, len(plan["band_edges"]) + 1))
                else:
                    table.is_dynamic(g)
                    colspec.append((g, table.radix(g)))
        weights, wgt = [], 1
        for _, radix in reversed(colspec):
            weights.append(wgt)
            wgt *= int(radix)
        weights = weights[::-1]
        n_keys = wgt
        if n_keys >= 2**31:
            raise ValueError("too many possible groups for dense accumulator slots")
        nk = 2 if bool(self.control) else 1
        nf = 2 if plan["flip"] else 1
        nctrl = self.CC.nshifts if do_control else 0
        # per-feature columns as plain numpy arrays, made once: a region's item then costs a few fancy-index takes instead
        # of a boolean-mask copy of the whole feature frame (the items are made while the GPU works on earlier regions)
        start_v, end_v = df["start"].values, df["end"].values
        chrom_codes, chrom_uniques = pd.factorize(df["chrom"])
        chrom_code = {str(c): i for i, c in enumerate(chrom_uniques)}
        center_v = np.ascontiguousarray(df["center"].values, dtype=np.float64)
        stbin_v = np.asarray(df["stBin"].values)
        region_rows = {name: (str(r["chrom"]), r["start"], r["end"]) for name, r in self.view_df.iterrows()}
        col_cache = {}

        def column(name):
            if name not in col_cache:
                col_cache[name] = df[name].to_numpy()
            return col_cache[name]

        def group_codes(g, source, idx):
            """Codes of group column ``g`` (values of feature column ``source``) for the features ``idx``: static
            dictionaries are applied to the whole column once."""
            if table.is_dynamic(g):
                return table.codes(g, column(source)[idx])
            key = ("codes", g)
            if key not in col_cache:
                col_cache[key] = np.asarray(table.codes(g, column(source)), dtype=np.int64)
            return col_cache[key][idx]

        def items():
            """One item per view region, made while the GPU works on the previous ones."""
            for ri, name in enumerate(region_names):
                yield make_item(ri, name)

        def make_item(ri, name):
            chrom, r_start, r_end = region_rows[name]
            idx = np.flatnonzero((chrom_codes == chrom_code.get(chrom, -1)) & (start_v >= r_start) & (end_v < r_end))
            center = center_v[idx]
            q, total = _native.pair_windows_count(center, self.CC.mindist, self.CC.maxdist)
            it = dict(index=ri, name=name, total=total, nctrl=nctrl, segs=(q[q > 0] * nctrl) if nctrl else np.zeros(0, np.int64),
                      owned=name in my_units and total > 0)
            if it["owned"]:
                lo_rel, hi_rel = self.view_df_extents[name]
                it.update(nb=hi_rel - lo_rel, ranges=my_units[name] or [(0, len(center))], center=center, per_offset=q,
                          stbin=np.ascontiguousarray(stbin_v[idx] - lo_rel, dtype=np.int32),
                          key1=None, key2=None, flipval=None, ident=None, band_weight=0)
                if plan["by_window"]:
                    mi = pd.MultiIndex.from_arrays([column("chrom")[idx], start_v[idx], end_v[idx]])
                    it["ident"] = np.ascontiguousarray(self._ident_index.reindex(mi).values, dtype=np.int32)
                else:
                    k1 = np.zeros(len(idx), dtype=np.int64)
                    k2 = np.zeros(len(idx), dtype=np.int64)
                    for (g, _), wt in zip(colspec, weights):
                        if g == "distance_band" and plan["band_edges"] is not None:
                            it["band_weight"] = wt
                            continue
                        codes = group_codes(g, g[:-1], idx)
                        if g[-1] == "1":
                            k1 += codes * wt
                        else:
                            k2 += codes * wt
                    it["key1"], it["key2"] = k1, k2
                if plan["flip"]:
                    if self.flip_negative_strand:
                        it["flipval"] = np.ascontiguousarray(column("strand")[idx] == "-", dtype=np.int32)
                    else:
                        fb = plan["flipby"]
                        it["flipval"] = np.ascontiguousarray(group_codes(fb + "1", fb, idx), dtype=np.int32)
            return it
        flags = 0
        if self.expected is True and self.ooe:
            flags |= _native.PUP_F_OOE
        if expctrl:
            flags |= _native.PUP_F_EXPCTRL
        if self.coverage_norm:
            flags |= _native.PUP_F_COVERAGE
        return dict(plan=plan, W=W, built=[], items=items, need_rng=bool(nctrl), colspec=colspec, table=table, first={}, nk=nk, nf=nf,
                    n_keys=n_keys, n_slots=n_keys * nk * nf, flags=flags, do_control=bool(self.control), expctrl=expctrl,
                    region_names=region_names, imbalance=imbalance)

    def _execute_device(self, job, acc, region_acc, exact, dist):
        """Device-window path: per region, control shifts (MT19937 replay) on a side stream, window generation on the
        prepare stream, pile-up on the compute stream.  Nothing here waits for the GPU."""
        import torch

        W, n_slots, flags, plan = job["W"], job["n_slots"], job["flags"], job["plan"]
        dev = torch.device("cuda", self._device)
        pipe = _native.make_pipeline(self._device, W, n_slots, flags)
        s_rng = pipe.s_side
        s_rng.wait_stream(torch.cuda.current_stream(dev))
        rng = _native.DeviceRng(self._device, stream=s_rng.cuda_stream) if job["need_rng"] else None
        first = torch.full((job["n_keys"],), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
        n_roi = torch.zeros(max(1, len(job["region_names"])), dtype=torch.int64, device=dev)
        pipe.s_prep.wait_stream(torch.cuda.current_stream(dev))
        edges = plan["band_edges"]
        edges = None if edges is None else np.ascontiguousarray(edges, dtype=np.float64)
        stride = _native.acc_stride(W)
        readies = []  # prepare-stream events of the regions submitted so far
        try:
            for it in job["items"]():
                dbin, shifts_ready = None, None
                if rng is not None and len(it["segs"]):
                    with torch.cuda.stream(s_rng):
                        if it["owned"]:
                            # the shift buffer of region k is reused by region k + 2: wait until region k's windows exist
                            if len(readies) >= 2:
                                s_rng.wait_event(readies[-2])
                            (dbin,) = pipe.scratch_i32("dbin", int(it["total"]) * it["nctrl"], stream=s_rng)
                        rng.control_shifts(it["segs"], self.CC.minshift, self.CC.maxshift, self.resolution, dbin,
                                           stream=s_rng.cuda_stream)
                        shifts_ready = s_rng.record_event()
                if not it["owned"]:
                    continue
                targets = 2 if it["ident"] is not None else 1
                # kept pairs per offset inside each of my row bands (anchor k in [k_lo, k_hi))
                bands = []
                for k_lo, k_hi in it["ranges"]:
                    if (k_lo, k_hi) == (0, len(it["center"])):
                        q_part = it["per_offset"]
                    else:
                        q_part, _ = _native.pair_windows_count(it["center"], self.CC.mindist, self.CC.maxdist, k_lo, k_hi)
                    bands.append((k_lo, k_hi, q_part, int(q_part.sum()) * (1 + it["nctrl"])))
                n_mine = sum(b[3] for b in bands)
                if n_mine == 0:
                    continue

                def generate(stream, it=it, dbin=dbin, shifts_ready=shifts_ready, n_mine=n_mine, targets=targets, bands=bands):
                    if shifts_ready is not None:
                        stream.wait_event(shifts_ready)
                    outs = pipe.scratch_i32("windows", n_mine * targets, count=3)
                    at = 0
                    for k_lo, k_hi, q_part, n_band in bands:
                        if n_band == 0:
                            continue
                        sub = tuple(o[at * targets : (at + n_band) * targets] for o in outs)
                        _native.pair_windows_device(
                            self._device, it["stbin"], it["center"], self.CC.mindist, self.CC.maxdist, it["nctrl"],
                            it["per_offset"], dbin, it["nb"], W, it["key1"], it["key2"], edges, it["band_weight"],
                            0 if it["flipval"] is None else (1 if self.flip_negative_strand else 2),
                            bool(plan["flip"] and plan["ignore_group_order"]), it["flipval"], it["ident"], job["nk"],
                            job["nf"], k_lo, k_hi, q_part, it["index"], sub[0], sub[1], sub[2], first_seen=first,
                            n_roi=n_roi[it["index"] : it["index"] + 1], stream=stream.cuda_stream)
                        at += n_band
                    return outs

                target = acc
                if exact:
                    target = region_acc[it["name"]] = _native.alloc_accumulator(n_slots * stride, self._device)
                readies.append(pipe.submit(self._region_kwargs(it["name"], flags), None, target, windows_on_device=generate))
            pipe.finish()
            s_rng.synchronize()
            if rng is not None:
                rng.store(stream=s_rng.cuda_stream)  # np.random continues where the reference's stream would be
        finally:
            if rng is not None:
                rng.close()
        if dist is not None and dist.world_size > 1:
            dist.all_reduce_min(first)
            dist.all_reduce(n_roi)
        fv = first.cpu().numpy()
        job["first"] = {int(k): (int(fv[k] >> 62), int((fv[k] >> 40) & 0xFFFFF), int(fv[k] & ((1 << 40) - 1)))
                        for k in np.nonzero(fv != np.iinfo(np.int64).max)[0]}
        counts = n_roi.cpu().numpy()
        for ri, name in enumerate(job["region_names"]):
            if counts[ri] > 0:
                logger.info(f"{(name, name)}: {int(counts[ri])}")
        return pipe

    def _execute_host(self, job, acc, region_acc, exact):
        """Host-window path: the window arrays of :meth:`_prepare` are uploaded region by region."""
        W, n_slots, flags = job["W"], job["n_slots"], job["flags"]
        stride = _native.acc_stride(W)
        pipe = _native.make_pipeline(self._device, W, n_slots, flags)
        for b in job["built"]:
            target = acc
            if exact:  # per-region accumulators: the reference's merge is not a plain sum when +inf occurs
                target = region_acc[b["name"]] = _native.alloc_accumulator(n_slots * stride, self._device)
            after = None
            if self.store_stripes:
                def after(region, stream, b=b):
                    # per-ROI centre row / column (coolpup.py:1164-1182); a by-window pair is computed once
                    sel = np.nonzero(b["valid"] & (b["rw"].kind == 0))[0]
                    hor, ver = region.stripes(np.ascontiguousarray(b["r0"][sel], dtype=np.int32),
                                              np.ascontiguousarray(b["c0"][sel], dtype=np.int32), W, stream=stream)
                    b["stripes"] = (sel, hor, ver)
            wins = (b["w_r0"], b["w_c0"], b["slot"])
            if self.rescale:
                wins += (b["w_h"], b["w_w"]) + ((b["mode"],) if job["expctrl"] else ())
            pipe.submit(self._region_kwargs(b["name"], flags), wins, target, after=after)
            n_roi = int(np.count_nonzero(b["valid"] & (b["rw"].kind == 0))) * b["targets"]
            if n_roi > 0:
                logger.info(f"{b['name'] if self.trans else (b['name'], b['name'])}: {n_roi}")
        pipe.finish()
        return pipe

    def _run(self, groupby, ignore_group_order, modify_2Dintervals_func, postprocess_func, regions=None,
             dist=None):
        """Accumulate all (or the given) view regions on the GPU; returns the merged ROI / control pile-ups."""
        import time

        _native.require_device()
        t0 = time.perf_counter()
        plan = self._plan(groupby, ignore_group_order, modify_2Dintervals_func, postprocess_func)
        plan["band_edges"] = self._band_edges(plan)
        on_device = self._device_windows_ok(plan)
        if on_device:
            job = self._prepare_device(plan, regions, dist)
        else:
            job = self._prepare(plan, regions, dist)
        t1 = time.perf_counter()
        W, n_slots = job["W"], job["n_slots"]
        stride = _native.acc_stride(W)
        acc = _native.alloc_accumulator(n_slots * stride, self._device)
        region_acc = {}
        execute = (lambda exact: self._execute_device(job, acc, region_acc, exact, dist)) if on_device else \
                  (lambda exact: self._execute_host(job, acc, region_acc, exact))
        pipe = execute(False)
        if dist is not None:
            dist.all_reduce(acc)
        # x / 0 = +inf can only arise with ooe and a zero expected value; when one really occurred (it poisons the sum),
        # the reference's merge quirk matters and the run is repeated with one accumulator per region (expected runs
        # draw no random numbers, so the repetition sees the same windows)
        exact = self._needs_exact_merge() and bool(_native.acc_has_inf(acc, W, n_slots))
        if exact:
            acc.zero_()
            pipe = execute(True)
            for a in region_acc.values():
                acc += a
            if dist is not None:
                dist.all_reduce(acc)
        stream = _native.current_stream(self._device)
        out, used = self._export_used(acc, job, stream)
        t2 = time.perf_counter()
        self._last_stats = {"windows": int(out["n"].sum()), "launches": pipe.launches, "regions": pipe.regions,
                            "imbalance": job["imbalance"], "n_slots": n_slots, "used_slots": int(len(used)),
                            "device_windows": bool(on_device), "host_prepare_s": t1 - t0, "gpu_phase_s": t2 - t1}
        _LAST_STATS.clear()
        _LAST_STATS.update(self._last_stats)
        grouped = bool(plan["groupby"]) or plan["by_window"]
        roi, ctrl = self._slots_to_pups(out, used, job, grouped)
        if exact:
            self._exact_merge(job, region_acc, roi, ctrl, grouped, dist, stream)
        if self.store_stripes:
            self._attach_stripes(job, roi, dist)
        return roi, ctrl

    def _export_used(self, acc, job, stream):
        """Decode the accumulator slots that hold at least one window (``n > 0``): with dense group keys most slots
        of a by-strand / by-distance / by-window run stay empty and are neither copied to the host nor decoded."""
        W, n_slots = job["W"], job["n_slots"]
        stride = _native.acc_stride(W)
        n_all = _native.acc_counts(acc, W, n_slots)
        used = np.nonzero(n_all > 0)[0]
        if len(used) == n_slots or n_slots <= 8:
            used = np.arange(n_slots)
            sub = acc
        else:
            import torch

            idx = torch.from_numpy(used).to(acc.device)
            sub = acc.view(n_slots, stride).index_select(0, idx).reshape(-1) if len(used) else acc[:0]
        out = _native.acc_export(sub, W, len(used), device=self._device, stream=stream, want_expected=job["expctrl"],
                                 want_cov=bool(self.coverage_norm)) if len(used) else {
            "sum": np.zeros((0, W, W)), "num": np.zeros((0, W, W), dtype=np.int64), "n": np.zeros(0, dtype=np.int64)}
        return out, used

    def _slot_pups(self, out, used, job):
        """{key: {kind: pup}} of the exported slots, flipped slots anti-transposed (coolpup.py:130) and merged."""
        nk, nf, W = job["nk"], job["nf"], job["W"]
        row_of = {int(s_): i for i, s_ in enumerate(used)}

        def antit(a):
            return a[..., ::-1, ::-1].swapaxes(-1, -2)

        def gather(field, key, kind):
            v = None
            for f in range(nf):
                i = row_of.get((key * nk + kind) * nf + f)
                if i is None:
                    continue
                w = out[field][i]
                if f == 1 and w.ndim == 2 and field in ("sum", "num"):
                    w = antit(w)
                v = w if v is None else v + w
            return v

        pups = {}
        for key in sorted({int(s_) // (nk * nf) for s_ in used}):
            per_kind = {}
            for kind in range(nk):
                n = gather("n", key, kind)
                if n is None or int(n) == 0:
                    continue
                p = {"data": gather("sum", key, kind), "num": gather("num", key, kind), "n": int(n),
                     "horizontal_stripe": [], "vertical_stripe": [], "coordinates": []}
                if "cov_start" in out:
                    p["cov_start"] = gather("cov_start", key, kind)
                    p["cov_end"] = gather("cov_end", key, kind)
                else:
                    p["cov_start"] = np.zeros(W)
                    p["cov_end"] = np.zeros(W)
                per_kind[kind] = p
            if job["expctrl"] and not self.rescale and 0 in per_kind:
                # bare expected blocks are Toeplitz, hence invariant under the anti-transpose flip
                p = per_kind[0]
                per_kind[1] = {"data": gather("exp_sum", key, 0), "num": gather("exp_num", key, 0), "n": p["n"],
                               "cov_start": p["cov_start"].copy(), "cov_end": p["cov_end"].copy(),
                               "horizontal_stripe": [], "vertical_stripe": [], "coordinates": []}
            if per_kind:
                pups[key] = per_kind
        return pups

    def _slots_to_pups(self, out, used, job, grouped):
        """Per-group ROI / control pile-ups in the reference's row order: groups by first valid ROI emission (regions
        in view order), ``"all"`` = sum over groups placed after the groups first seen in the first view region
        (coolpup.py:1272-1275, 1511-1520)."""
        W = job["W"]
        first = job["first"]
        pups = self._slot_pups(out, used, job)
        order = sorted((k for k in pups if k in first), key=lambda k: first[k])
        roi, ctrl = {}, {}
        for key in order:
            name = self._decode_key(job, key)
            if 0 in pups[key]:
                roi[name] = pups[key][0]
            if 1 in pups[key]:
                ctrl[name] = pups[key][1]
        has_ctrl = job["do_control"] or job["expctrl"]
        if not grouped:
            roi.setdefault("all", _empty_pup(W))
            if has_ctrl:
                ctrl.setdefault("all", _empty_pup(W))
            return roi, ctrl
        first_region = min((v[1] for v in first.values()), default=0)
        n_first_region = sum(1 for k in order if 0 in pups[k] and first[k][0] == 0 and first[k][1] == first_region)
        for d, present in ((roi, True), (ctrl, has_ctrl)):
            if not present:
                continue
            tot = _empty_pup(W)
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

    def _exact_merge(self, job, region_acc, roi, ctrl, grouped, dist, stream):
        """Replace ``data`` of the merged pile-ups by what the reference's reduce produces when +inf pixels exist.

        Reference semantics (coolpup.py:1272-1282, 1511-1531; lib/puputils.py:88-113): inside a region nansum keeps
        +inf; building the region's ``"all"`` from its groups rebinds every group's ``data`` to ``nan_to_num(data)``
        (+inf -> 1.797e308) and sums them; merging regions applies ``nan_to_num`` again whenever a key occurs in at
        least two regions (a single occurrence is taken as it is).  1.797e308 + 1.797e308 overflows back to +inf, which
        the final ``== inf -> NaN`` rule removes.  Every other field is additive and comes from the summed accumulator.
        """
        W = job["W"]
        per_region = {}
        for name, a in region_acc.items():
            out, used = self._export_used(a, job, stream)
            pups = self._slot_pups(out, used, job)
            per_region[name] = {kind: {self._decode_key(job, k): v[kind]["data"] for k, v in pups.items() if kind in v}
                                for kind in (0, 1)}
        if dist is not None and dist.world_size > 1:
            merged = {}
            for part in dist.all_gather_object(per_region):
                for name, kinds in part.items():
                    tgt = merged.setdefault(name, {0: {}, 1: {}})
                    for kind, groups in kinds.items():
                        for g, data in groups.items():  # a region split over ranks: nansum semantics, inf kept
                            tgt[kind][g] = data if g not in tgt[kind] else tgt[kind][g] + data
            per_region = merged
        has_ctrl = job["do_control"] or job["expctrl"]
        for kind, d in ((0, roi), (1, ctrl)):
            if kind == 1 and not has_ctrl:
                continue
            contrib = {}  # key -> [data per region, view order]
            for name in job["region_names"]:  # every view region the reference maps pileup_region over
                groups = dict(per_region.get(name, {0: {}, 1: {}})[kind])
                if grouped:
                    tot = np.zeros((W, W))
                    for g in groups:
                        groups[g] = np.nan_to_num(groups[g])
                        tot = np.nan_to_num(tot) + groups[g]
                    groups["all"] = tot
                elif "all" not in groups:
                    groups["all"] = np.zeros((W, W))
                for g, data in groups.items():
                    contrib.setdefault(g, []).append(data)
            for g, p in d.items():
                lst = contrib.get(g, [])
                if len(lst) == 1:
                    p["data"] = lst[0]
                elif lst:
                    total = lst[0]
                    for x in lst[1:]:
                        total = np.nan_to_num(total) + np.nan_to_num(x)
                    p["data"] = total

    def _attach_stripes(self, job, roi, dist=None):
        """Per-group lists of stripes / coordinates in the reference's order: regions in view order; within a region
        the groups in order of first appearance, each in stream order; "all" concatenates the region's groups
        (sum_pups list concatenation, lib/puputils.py:105-107; coolpup.py:1272-1275)."""
        grouped = bool(job["plan"]["groupby"]) or job["plan"]["by_window"]
        per_region = {}
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
            keys = b["key"].reshape(len(rw), -1)[sel]  # [n_sel, targets]
            per_group = {}
            order = []
            for j in range(len(sel)):
                for k in keys[j]:
                    name = self._decode_key(job, int(k)) if grouped else "all"
                    if name not in per_group:
                        per_group[name] = []
                        order.append(name)
                    per_group[name].append(j)
            per_region[b["index"]] = [(name, [(hor[j], ver[j], coords[j]) for j in per_group[name]]) for name in order]
        if dist is not None and dist.world_size > 1:  # whole regions per rank (stripes are never window-split)
            merged = {}
            for part in dist.all_gather_object(per_region):
                merged.update(part)
            per_region = merged
        lists = {k: {"horizontal_stripe": [], "vertical_stripe": [], "coordinates": []} for k in roi}
        for index in sorted(per_region):
            for name, rows in per_region[index]:
                if name not in lists:
                    continue
                for h, v, c in rows:
                    for dst in ([name, "all"] if grouped else [name]):
                        lists[dst]["horizontal_stripe"].append(h)
                        lists[dst]["vertical_stripe"].append(v)
                        lists[dst]["coordinates"].append(c)
        for k, p in roi.items():
            p.update(lists[k])

    def _feature_ident(self, rw):
        """Global integer identity of each feature of the region table for by-window grouping."""
        if not hasattr(self, "_ident_index"):
            df = self.CC.intervals
            triples = pd.MultiIndex.from_arrays([df["chrom"].values, df["start"].values, df["end"].values])
            uniq = triples.unique()
            self._ident_values = list(uniq)
            self._ident_index = pd.Series(np.arange(len(uniq), dtype=np.int64), index=uniq)
        if rw is None:
            return None

        def of(tab):
            mi = pd.MultiIndex.from_arrays([tab["chrom"].values, tab["start"].values, tab["end"].values])
            return self._ident_index.reindex(mi).values.astype(np.int64)

        return of(rw.sel), (of(rw.sel2) if rw.sel2 is not rw.sel else of(rw.sel))

    def pileup_region(self, region1, region2=None, groupby=[], modify_2Dintervals_func=None, postprocess_func=None,
                      extra_sum_funcs=None):
        """Accumulated pile-ups of one view region: ``{"ROI": {group: pup}, "control": {...}}`` (coolpup.py:1285-1358)."""
        if extra_sum_funcs:
            raise NotImplementedError("extra_sum_funcs callbacks are not supported by the B200 path")
        if self.trans != (region2 is not None and region2 != region1):
            raise ValueError("pileup_region: two different regions are needed exactly when trans=True")
        roi, ctrl = self._run(groupby, self.ignore_group_order, modify_2Dintervals_func, postprocess_func,
                              regions=[(region1, region2) if self.trans else region1])
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
            W = self._out_size()
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
        if self.trans:
            raise ValueError("Cannot do by-distance pileups for trans")
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
        if self.trans:
            raise ValueError("Cannot do by-distance pileups for trans")
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


def _empty_pup(W):
    return {"data": np.zeros((W, W)), "num": np.zeros((W, W), dtype=np.int64), "n": 0, "cov_start": np.zeros(W),
            "cov_end": np.zeros(W), "horizontal_stripe": [], "vertical_stripe": [], "coordinates": []}


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
        need = ["region1", "region2"] + ([] if trans else ["dist"]) + [expected_value_col]
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

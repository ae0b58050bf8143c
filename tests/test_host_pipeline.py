"""CPU tests of the product's host side (window generation, slots, flips, final normalisation, DataFrame schema)
against the real reference's golden vectors.  The CUDA entry points are replaced by tests/emulator.py, a numpy
mirror of the kernels' arithmetic and accumulator layout; the real kernels are checked in test_gpu_parity.py."""
import warnings

import numpy as np
import pytest

import emulator
import golden_util as gu
from oracle.pileup_oracle import key_repr

CASES = [n for n in gu.all_cases() if gu.manifest()[n]["windows"] <= 4000]


@pytest.fixture()
def emu(monkeypatch):
    emulator.install(monkeypatch)


@pytest.mark.parametrize("name", CASES)
def test_pileup_api_matches_reference_golden(emu, name):
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pups = cp.pileup(clr, feats, **kw)
    z, _ = gu.load_golden(name)
    if "group" in pups.columns:
        keys = [key_repr(g) for g in pups["group"]]
    else:
        keys = [repr((r.chrom, int(r.start), int(r.end))) for r in pups.itertuples()]
    assert keys == [str(k) for k in z["row_keys"]]
    assert list(pups.columns) == [str(c) for c in z["columns"]]
    for i in range(len(keys)):
        g = {f.split(".", 1)[1]: z[f] for f in z.files if f.startswith(f"row{i}.")}
        a, b = np.asarray(pups["data"].iloc[i], dtype=float), g["data"]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = np.isfinite(b)
        np.testing.assert_allclose(a[m], b[m], rtol=1e-9)
        assert int(pups["n"].iloc[i]) == int(g["n"])
        assert np.array_equal(np.asarray(pups["num"].iloc[i]), g["num"])
        if "control_n" in g:
            assert int(pups["control_n"].iloc[i]) == int(g["control_n"])
            assert np.array_equal(np.asarray(pups["control_num"].iloc[i]), g["control_num"])
        if "vertical_stripe" in g:
            for f in ("vertical_stripe", "horizontal_stripe"):
                np.testing.assert_allclose(np.asarray(pups[f].iloc[i], dtype=float), g[f], rtol=1e-9, equal_nan=True)
            assert np.array_equal(np.asarray(pups["coordinates"].iloc[i]).astype(str), g["coordinates"])


@pytest.mark.parametrize("name", ["toy_controls", "scc1_loops_ctrl", "scc1_ctcf_pairs_strand_dist", "scc1_ctcf_pairs_arms",
                                  "toy_local_raw", "scc1_loops_dist"])
def test_window_arrays_match_reference_stream(name):
    """CoordCreator.region_windows == the window stream the reference's pos_stream emitted (order, RNG draws)."""
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs(name)
    z, _ = gu.load_golden(name)
    seed, nshifts = kw.get("seed"), kw.get("nshifts", 0)
    view = kw.get("view_df")
    view = cp.make_cooler_view(clr) if view is None else cp.make_viewframe(view)
    if seed is not None:
        np.random.seed(seed)
    cc = cp.CoordCreator(feats, clr.binsize, features_format=kw.get("features_format", "bed"), flank=kw.get("flank", 100000),
                         chroms=list(view["chrom"].unique()), nshifts=nshifts, mindist=kw.get("mindist", "auto"),
                         maxdist=kw.get("maxdist"), local=kw.get("local", False), subset=kw.get("subset", 0), seed=seed)
    view = view[view["chrom"].isin(cc.final_chroms)]
    for _, r in view.iterrows():
        rw = cc.region_windows((r["chrom"], r["start"], r["end"]), control=nshifts > 0)
        name_r = r["name"]
        assert np.array_equal(rw.st1, z[f"win.{name_r}.st1"])
        assert np.array_equal(rw.st2, z[f"win.{name_r}.st2"])
        assert np.array_equal(rw.kind, z[f"win.{name_r}.kind"])


def test_unsupported_features_raise():
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs("toy_strand_balanced")
    with pytest.raises(ValueError):
        cp.pileup(clr, feats, features_format="bed", flank=2_000_000, trans=True, by_distance=True, view_df=kw["view_df"])
    with pytest.raises(ValueError):
        cp.pileup(clr, feats, features_format="bed", flank=2_000_000, trans=True, local=True, view_df=kw["view_df"])
    with pytest.raises(NotImplementedError):
        cp.pileup(clr, feats, features_format="bed", flank=2_000_000, rescale=True, rescale_flank=1, store_stripes=True)
    with pytest.raises(ValueError):
        cp.pileup(clr, feats, features_format="bed", flank=2_000_000, rescale=True, rescale_flank=1, rescale_size=10)
    with pytest.raises(ValueError):
        cp.pileup(clr, feats, features_format="bed", flank=2_000_000, local=True, by_distance=True)
    with pytest.raises(ValueError):
        cp.pileup(clr, feats, features_format="bed", flank=2_500_000, mindist=0)  # flank not a multiple of the bin size


def test_custom_modify_func_and_extra_groupby(emu):
    """User DataFrame callbacks (modify_2Dintervals_func) and arbitrary groupby columns go through the slow frame path."""
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs("toy_strand_balanced")
    feats = feats.copy()
    feats["cls"] = ["a", "b", "a", "b", "a", "b"]
    view = kw["view_df"]
    cc = cp.CoordCreator(feats, clr.binsize, features_format="bed", flank=2_000_000, mindist=0, chroms=["chr1", "chr2"])
    pu = cp.PileUpper(clr, cc, view_df=view)

    def tag(df):
        df["far"] = df["distance"] > 4_000_000
        return df

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pu.pileupsWithControl(groupby=["cls1", "far"], modify_2Dintervals_func=tag)
        base = pu.pileupsWithControl()
    assert int(out.loc[out["group"] == "all", "n"].iloc[0]) == 6
    assert sorted(int(n) for n in out.loc[out["group"] != "all", "n"]) == sorted(
        [int(x) for x in np.bincount([0, 0, 1, 1, 2, 2])] ) or out["n"].sum() == 12
    a = out.loc[out["group"] == "all", "data"].iloc[0]
    b = base.loc[base["group"] == "all", "data"].iloc[0]
    np.testing.assert_allclose(np.nan_to_num(a), np.nan_to_num(b), rtol=1e-12)


def test_missing_group_values_form_their_own_group(emu):
    """A NaN in a groupby column must neither wrap around to the last group nor crash: the windows of
    that feature land in a (nan, ...) group, every other count is unchanged."""
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs("toy_strand_balanced")
    feats = feats.copy()
    feats["strand"] = feats["strand"].astype(object)
    feats.loc[0, "strand"] = np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cp.pileup(clr, feats, **kw)
        base = cp.pileup(clr, gu.case_inputs("toy_strand_balanced")[1], **kw)
    tot = int(out.loc[out["group"] == "all", "n"].iloc[0])
    assert tot == int(base.loc[base["group"] == "all", "n"].iloc[0]) == 6
    groups = [g for g in out["group"] if not isinstance(g, str)]
    nan_groups = [g for g in groups if any(isinstance(x, float) and x != x for x in g)]
    assert nan_groups, groups
    assert sum(int(n) for g, n in zip(out["group"], out["n"]) if not isinstance(g, str)) == 6


def test_dynamic_group_dictionaries_are_merged_over_ranks():
    """Columns made by a user callback get rank-local codes while the regions are laid out; finalize() must map
    them onto ONE dictionary (the same integer meant different values on different ranks)."""
    from coolpuppy_b200.coolpup import _GroupTable

    class CC:
        kind = "bed"
        import pandas as pd

        intervals = pd.DataFrame({"chrom": ["chr1"], "start": [0], "end": [1]})

    class FakeDist:
        world_size = 2

        def __init__(self, other):
            self.other = other

        def all_gather_object(self, obj):
            return [obj, self.other]

    t0, t1 = _GroupTable(CC), _GroupTable(CC)
    c0 = t0.codes("cls", np.array(["b", "a", "b", None], dtype=object))
    c1 = t1.codes("cls", np.array(["c", "a"], dtype=object))
    assert list(c0) == [0, 1, 0, 2] and list(c1) == [0, 1]
    t0.finalize(FakeDist(t1._dynamic["cls"]["vals"]))
    t1.finalize(FakeDist(t0._dynamic["cls"]["vals"]))
    assert t0.radix("cls") == t1.radix("cls") == 4
    v0 = [t0.value("cls", k) for k in t0.remap("cls", c0)]
    v1 = [t1.value("cls", k) for k in t1.remap("cls", c1)]
    assert v0 == ["b", "a", "b", None] and v1 == ["c", "a"]
    assert [t0.value("cls", k) for k in range(4)] == [t1.value("cls", k) for k in range(4)] == ["a", "b", "c", None]


class _AllPartsDist:
    """A fake sharder that hands THIS rank every part of every region as separate units (collectives are identities):
    a rank can own several parts of one region (round-2 bug: the later part used to replace the earlier one)."""

    world_size = 2
    rank = 0

    def my_units(self, items, costs, max_share=0.25):
        units = []
        for j, it in enumerate(items):
            parts = 3 if j % 2 == 0 else 2
            units += [(it, p, parts) for p in range(parts)]
        return units, 1.0

    def all_reduce(self, t):
        return t

    def all_reduce_min(self, t):
        return t

    def merge_min(self, m):
        return m

    def all_gather_object(self, o):
        return [o]


@pytest.mark.parametrize("name", ["toy_strand_dist_ctrl", "toy_bywindow", "scc1_ctcf_pairs_arms", "scc1_loops_ctrl",
                                  "toy_local_raw", "toy_zero_expected_strand"])
def test_several_parts_of_one_region_on_one_rank(emu, name):
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pups = cp.pileup(clr, feats, dist=_AllPartsDist(), **kw)
    z, _ = gu.load_golden(name)
    for i in range(len(pups)):
        g = {f.split(".", 1)[1]: z[f] for f in z.files if f.startswith(f"row{i}.")}
        a, b = np.asarray(pups["data"].iloc[i], dtype=float), g["data"]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = np.isfinite(b)
        np.testing.assert_allclose(a[m], b[m], rtol=1e-9)
        assert int(pups["n"].iloc[i]) == int(g["n"]) and np.array_equal(np.asarray(pups["num"].iloc[i]), g["num"])


def test_part_ranges_cover_features_once():
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs("scc1_ctcf_pairs_arms")
    view = kw["view_df"]
    cc = cp.CoordCreator(feats, clr.binsize, features_format="bed", flank=50_000, mindist=0, maxdist=2_000_000,
                         chroms=list(view["chrom"].unique()))
    pu = cp.PileUpper(clr, cc, view_df=view, clr_weight_name=None)
    for name in pu.view_df.index:
        n = len(pu._feature_costs(name))
        for parts in (2, 5):
            got = []
            for p in range(parts):
                got += [k for lo, hi in pu._part_ranges(name, [p], parts) for k in range(lo, hi)]
            assert sorted(got) == list(range(n))
            merged = pu._part_ranges(name, range(parts), parts)
            assert merged == [(0, n)]
        c = pu._feature_costs(name)
        cuts = [pu._part_ranges(name, [p], 4) for p in range(4)]
        loads = [sum(c[lo:hi].sum() for lo, hi in r) for r in cuts]
        assert max(loads) <= 0.5 * c.sum()  # equal-cost cut points, up to one feature's cost


def test_sharding_cost_model_follows_the_two_kernels():
    """PileUpper._feature_costs: the Poisson occupancy fit recovers the scale of a synthetic matrix, and a window costs
    the dense-band kernel's flat price where the matrix is dense, the sparse kernel's per-pixel price further out."""
    from coolpuppy_b200.coolpup import PileUpper

    nb = 20000
    s_ = np.arange(1, nb)
    nnz = float(np.sum((nb - s_) * -np.expm1(-300.0 / s_))) + nb
    lam = PileUpper._fit_poisson_scale(nb, nnz)
    assert abs(lam - 300.0) / 300.0 < 0.02
    W = 83
    dense = PileUpper._COST_DENSE_PER_CELL * W * W
    fill_far = -np.expm1(-lam / 10000.0)
    sparse_far = PileUpper._COST_SPARSE_PER_ROW * W + PileUpper._COST_SPARSE_PER_PIXEL * W * W * fill_far
    assert sparse_far < dense  # far windows are cheaper through the sparse kernel ...
    fill_near = -np.expm1(-lam / 200.0)
    sparse_near = PileUpper._COST_SPARSE_PER_ROW * W + PileUpper._COST_SPARSE_PER_PIXEL * W * W * fill_near
    assert fill_near > PileUpper._DENSE_FILL and sparse_near > dense  # ... near ones through the dense band

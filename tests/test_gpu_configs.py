"""GPU parity of the BASELINE.json configurations that are not the bench line, at sizes the oracle finishes in
seconds, through the public Python API on a synthetic cooler:
  configs[2]: bedpe loops, pad = 41, balanced, expected ooe
  configs[4]: all-vs-all stranded sites, by_strand + by_distance, pad = 101 (W = 203: the tile is split in row bands)
"""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _compare(pups, ref):
    from oracle.pileup_oracle import key_repr

    # the wrappers re-sort the rows (coolpup.py:1826-1832, 1910-1918; pinned by tests/golden): compare by group key
    by_key = ref.by_key()
    assert sorted(key_repr(g) for g in pups["group"]) == sorted(by_key)
    for _, row in pups.iterrows():
        o = by_key[key_repr(row["group"])]
        assert int(row["n"]) == int(o["n"])
        assert np.array_equal(np.asarray(row["num"]), np.asarray(o["num"]))
        a, b = np.asarray(row["data"], dtype=float), np.asarray(o["data"], dtype=float)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = np.isfinite(b)
        np.testing.assert_allclose(a[m], b[m], rtol=RTOL)
        if "control_n" in o:
            assert int(row["control_n"]) == int(o["control_n"])
            assert np.array_equal(np.asarray(row["control_num"]), np.asarray(o["control_num"]))


@pytest.fixture(scope="module")
def genome():
    from coolpuppy_b200 import _native
    from coolpuppy_b200.synthetic import synthetic_cooler

    if _native.device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    sizes = {"chr1": 52_000_000, "chr2": 38_000_000, "chr3": 9_000_000}
    return synthetic_cooler(sizes, binsize=10_000, depth=60.0, seed=77, device="cuda"), sizes


def test_config2_bedpe_loops_pad41_ooe(genome):
    from coolpuppy_b200 import coolpup as cp
    from coolpuppy_b200.synthetic import synthetic_loops
    from oracle.pileup_oracle import oracle_pileup

    (clr, exp), sizes = genome
    loops = synthetic_loops(1500, chromsizes=sizes, binsize=10_000, flank=410_000, seed=5, dmin=900_000, dmax=8_000_000)
    kw = dict(features_format="bedpe", flank=410_000, expected_df=exp, ooe=True, clr_weight_name="weight")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pups = cp.pileup(clr, loops, **kw)
        ref = oracle_pileup(clr, loops, **kw)
    assert int(pups["n"].iloc[0]) > 1000
    _compare(pups, ref)


def test_config4_strand_distance_pad101(genome):
    from coolpuppy_b200 import coolpup as cp
    from coolpuppy_b200.synthetic import synthetic_sites
    from oracle.pileup_oracle import oracle_pileup

    (clr, exp), sizes = genome
    sites, npairs = synthetic_sites(1500, chromsizes=sizes, binsize=10_000, flank=1_010_000, seed=3)
    kw = dict(features_format="bed", flank=1_010_000, by_strand=True, by_distance=True, expected_df=exp, ooe=True,
              clr_weight_name="weight")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pups = cp.pileup(clr, sites, **kw)
        ref = oracle_pileup(clr, sites, **kw)
    assert pups["data"].iloc[0].shape == (203, 203) and len(pups) > 5
    _compare(pups, ref)


def test_config3_shape_controls_allpairs_pad41(genome):
    """The bench configuration (all-vs-all pairs, nshifts controls, balanced, no expected) at oracle-checkable size."""
    from coolpuppy_b200 import coolpup as cp
    from coolpuppy_b200.synthetic import synthetic_sites
    from oracle.pileup_oracle import oracle_pileup

    (clr, exp), sizes = genome
    sites, npairs = synthetic_sites(400, chromsizes=sizes, binsize=10_000, flank=410_000, seed=9)
    kw = dict(features_format="bed", flank=410_000, nshifts=3, seed=0, clr_weight_name="weight")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pups = cp.pileup(clr, sites, **kw)
        ref = oracle_pileup(clr, sites, **kw)
    _compare(pups, ref)


@pytest.mark.parametrize("fixture,weight", [("CN.mm9.1000kb.cool", "weight"), ("CN.mm9.1000kb.cool", None),
                                            ("Scc1-control.10000.cool", None)])
def test_expected_cis_on_gpu_matches_numpy(fixture, weight):
    """pup_expected_cis (per-diagonal sums + valid-pair counts on the device) == the numpy restatement of
    cooltools expected-cis that the golden cases use."""
    import os

    import pandas as pd

    from coolpuppy_b200 import _native
    from coolpuppy_b200.coolio import Cooler
    from coolpuppy_b200.expected import expected_cis, expected_cis_gpu

    if _native.device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    clr = Cooler(os.path.join(os.path.dirname(__file__), "fixtures", fixture))
    names = list(clr.chromnames)[:4]
    view = pd.DataFrame({"chrom": names, "start": 0, "end": [int(clr.chromsizes[c]) for c in names], "name": names})
    # an arm-like sub-region too: pixels that leave the region must be ignored
    view.loc[len(view)] = [names[0], 0, int(clr.chromsizes[names[0]]) // 2 // clr.binsize * clr.binsize, "half"]
    a = expected_cis(clr, view_df=view, clr_weight_name=weight, ignore_diags=2)
    b = expected_cis_gpu(clr, view_df=view, clr_weight_name=weight, ignore_diags=2)
    assert list(a.columns) == list(b.columns) and len(a) == len(b)
    assert np.array_equal(a["n_valid"].values, b["n_valid"].values)
    for c in a.columns:
        if c.endswith((".sum", ".avg")):
            np.testing.assert_allclose(b[c].values, a[c].values, rtol=1e-10, equal_nan=True)


def test_expected_cis_gpu_on_synthetic_genome(genome):
    from coolpuppy_b200.expected import expected_cis, expected_cis_gpu

    (clr, exp), sizes = genome
    a = expected_cis(clr, clr_weight_name="weight", ignore_diags=2)
    b = expected_cis_gpu(clr, clr_weight_name="weight", ignore_diags=2)
    assert np.array_equal(a["n_valid"].values, b["n_valid"].values)
    np.testing.assert_allclose(b["balanced.avg"].values, a["balanced.avg"].values, rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(b["count.sum"].values, a["count.sum"].values, rtol=1e-10, equal_nan=True)


def test_by_window_with_thousands_of_features(genome):
    """By-window pile-up (one group per feature, every window goes to both anchors' groups) with > 2000 features --
    thousands of accumulator slots, most of them sparse -- against the oracle, on both window paths."""
    import os

    from coolpuppy_b200 import coolpup as cp
    from oracle.pileup_oracle import key_repr, oracle_pileup
    import pandas as pd

    (clr, exp), sizes = genome
    rng = np.random.default_rng(21)
    rows = []
    for c, L in sizes.items():
        n = int(2300 * L / sum(sizes.values()))
        pos = np.sort(rng.choice(np.arange(60_000, L - 70_000, 10_000), size=n, replace=False)) + rng.integers(0, 9_000, n)
        rows.append(pd.DataFrame({"chrom": c, "start": pos, "end": pos + 500}))
    feats = pd.concat(rows, ignore_index=True)
    assert len(feats) > 2000
    kw = dict(features_format="bed", flank=20_000, by_window=True, mindist=60_000, maxdist=130_000, clr_weight_name="weight")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = oracle_pileup(clr, feats, **kw).by_key()
        for mode in ("1", "0"):
            os.environ["PUP_DEVICE_WINDOWS"] = mode
            try:
                pups = cp.pileup(clr, feats, **kw)
            finally:
                os.environ.pop("PUP_DEVICE_WINDOWS", None)
            assert cp._LAST_STATS["device_windows"] == (mode == "1")
            keys = [repr((r.chrom, int(r.start), int(r.end))) if r.chrom != "all" else "all" for r in pups.itertuples()]
            assert len(keys) > 2000 and set(keys) == set(ref)
            for k, (_, row) in zip(keys, pups.iterrows()):
                o = ref[k]
                assert int(row["n"]) == int(o["n"])
                assert np.array_equal(np.asarray(row["num"]), np.asarray(o["num"]))
                a, b = np.asarray(row["data"], dtype=float), np.asarray(o["data"], dtype=float)
                assert np.array_equal(np.isnan(a), np.isnan(b))
                m = np.isfinite(b)
                np.testing.assert_allclose(a[m], b[m], rtol=RTOL)

"""CPU tests: .cool reader, oracle vs the real reference's golden vectors, reference-asserted known answers."""
import os
import warnings

import numpy as np
import pytest

import golden_util as gu
from coolpuppy_b200.coolio import Cooler, MemCooler
from oracle.pileup_oracle import key_repr, oracle_pileup

FAST_CASES = [n for n in gu.all_cases() if gu.manifest()[n]["windows"] <= 4000]
SLOW_CASES = [n for n in gu.all_cases() if n not in FAST_CASES]


def test_hdf5_reader_known_answer(fixtures_dir):
    """Raw counts of chr1 rows 100-104 x cols 103-107 and the stripe vector asserted by the reference
    (tests/test_coolpup.py:167-172)."""
    clr = Cooler(os.path.join(fixtures_dir, "CN.mm9.1000kb.cool"))
    assert clr.binsize == 1_000_000 and len(clr.chromnames) == 22
    m = clr.matrix(sparse=True, balance=False).fetch("chr1").tocsr()
    block = m[100:105, 103:108].toarray()
    expect = np.array([[8523, 9432, 9998, 5024, 4276], [13450, 13043, 11376, 5960, 5220],
                       [20362, 14680, 10852, 5373, 4535], [48005, 20683, 11287, 4844, 3908],
                       [20683, 59415, 22015, 6460, 4782]])
    assert np.array_equal(block, expect)
    assert list(block[:, 2][::-1]) == [22015, 11287, 10852, 11376, 9998]
    assert (m != m.T).nnz == 0
    assert clr.extent(("chr2", 100_000_000, 150_000_000)) == (298, 348)
    w = clr.bins()["weight"].fetch("chr1").values
    assert np.isnan(w).sum() > 0


def test_region_csr_matches_matrix_fetch(fixtures_dir):
    import scipy.sparse as sp

    clr = Cooler(os.path.join(fixtures_dir, "Scc1-control.10000.cool"))
    lo, hi = clr.extent("chr19")
    ip, col, cnt = clr.region_csr(lo, hi)
    m = clr.matrix(sparse=True, balance=False).fetch("chr19").tocsr()
    m2 = sp.csr_matrix((cnt, col, ip), shape=(hi - lo, hi - lo))
    assert abs(m - m2).sum() == 0
    assert all(np.all(np.diff(col[ip[i] : ip[i + 1]]) > 0) for i in range(0, hi - lo, 97))


def test_region_upper_csr_sorts_out_of_order_rows(fixtures_dir):
    """CN.mm9.1000kb.cool (the reference's own fixture) stores some pixels out of (bin1, bin2) order; the device
    mirror step needs sorted rows, so the reader must deliver them sorted."""
    clr = Cooler(os.path.join(fixtures_dir, "CN.mm9.1000kb.cool"))
    assert not np.all(np.diff(clr._bin1 * 10000 + clr._bin2) > 0)  # the file really is out of order
    for chrom in ("chr1", "chr2", "chr7"):
        lo, hi = clr.extent(chrom)
        ip, col, cnt = clr.region_upper_csr(lo, hi)
        rows = np.repeat(np.arange(hi - lo), np.diff(ip))
        same = rows[1:] == rows[:-1]
        assert np.all(np.diff(col.astype(np.int64))[same] > 0)
        m = clr.matrix(sparse=True, balance=False).fetch(chrom).tocsr()
        keep = col < hi - lo
        import scipy.sparse as sp

        up = sp.csr_matrix((cnt[keep], (rows[keep], col[keep])), shape=(hi - lo, hi - lo))
        assert abs(sp.triu(m) - up).sum() == 0


def test_memcooler_roundtrip():
    rng = np.random.default_rng(0)
    sizes = {"a": 1050, "b": 730}
    nb = 11 + 8
    i, j = np.triu_indices(nb)
    keep = rng.random(i.size) < 0.4
    i, j = i[keep], j[keep]
    cnt = rng.integers(1, 9, i.size).astype(np.int32)
    w = rng.random(nb)
    w[3] = np.nan
    clr = MemCooler(sizes, 100, i, j, cnt, {"weight": w})
    assert clr.extent("b") == (11, 19)
    dense = np.zeros((nb, nb))
    dense[i, j] = cnt
    dense[j, i] = cnt
    got = clr.matrix(sparse=True, balance=False).fetch("b").toarray()
    assert np.array_equal(got, dense[11:, 11:])
    bal = clr.matrix(sparse=True, balance="weight").fetch("a").toarray()
    ref = dense[:11, :11] * np.outer(w[:11], w[:11])
    m = dense[:11, :11] > 0
    np.testing.assert_allclose(bal[m & ~np.isnan(ref)], ref[m & ~np.isnan(ref)])


def _check_oracle_case(name):
    clr, feats, kw = gu.case_inputs(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = oracle_pileup(clr, feats, **kw)
    z, rows = gu.load_golden(name)
    bk = res.by_key()
    # the by-window wrapper renames the "all" row to ('all', -1, -1) (coolpup.py:1744-1746)
    rows = {("all" if k.startswith("('all'") else k): v for k, v in rows.items()}
    assert set(bk) == set(rows)
    for k, g in rows.items():
        o = bk[k]
        a, b = np.asarray(o["data"], dtype=float), g["data"]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = np.isfinite(b)
        np.testing.assert_allclose(a[m], b[m], rtol=1e-12, atol=0)  # observed: bit-identical
        assert int(o["n"]) == int(g["n"])
        assert np.array_equal(np.asarray(o["num"]), g["num"])
        if "control_n" in g:
            assert int(o["control_n"]) == int(g["control_n"])
            assert np.array_equal(np.asarray(o["control_num"]), g["control_num"])
        if "vertical_stripe" in g:
            np.testing.assert_allclose(np.asarray(o["vertical_stripe"], dtype=float), g["vertical_stripe"], equal_nan=True)
            np.testing.assert_allclose(np.asarray(o["horizontal_stripe"], dtype=float), g["horizontal_stripe"], equal_nan=True)
    # window streams (pair order, distance filter, np.random control-shift order) and per-region accumulators
    trans_key = {k[0]: k for k in res.windows if isinstance(k, tuple)}  # trans: the golden is keyed by region1
    for r in z["regions"]:
        r = str(r)
        w = res.windows.get(trans_key.get(r, r))
        if w is None:
            assert len(z[f"win.{r}.st1"]) == 0
            continue
        assert np.array_equal(w["st1"], z[f"win.{r}.st1"]) and np.array_equal(w["st2"], z[f"win.{r}.st2"])
        assert np.array_equal(w["kind"], z[f"win.{r}.kind"])
        gk = z[f"win.{r}.group_keys"]
        assert [str(gk[i]) for i in z[f"win.{r}.group"]] == w["group"]
        acc_keys = [str(k) for k in z[f"acc.{r}.keys"]]
        for j, kk in enumerate(acc_keys):
            kind, key = kk.split("|", 1)
            mine = res.regions[trans_key.get(r, r)][(kind, key)]
            assert int(mine["n"]) == int(z[f"acc.{r}.{j}.n"])
            assert np.array_equal(np.asarray(mine["num"]), z[f"acc.{r}.{j}.num"])
            np.testing.assert_allclose(np.nan_to_num(np.asarray(mine["data"], dtype=float)),
                                       np.nan_to_num(z[f"acc.{r}.{j}.data"]), rtol=1e-12)


@pytest.mark.parametrize("name", FAST_CASES)
def test_oracle_matches_reference_golden(name):
    _check_oracle_case(name)


@pytest.mark.parametrize("name", ["scc1_loops_ctrl", "scc1_ctcf_local_ooe"])
def test_oracle_matches_reference_golden_large(name):
    _check_oracle_case(name)


def test_reference_asserted_counts():
    """The n tables the reference's own tests assert (tests/test_coolpup.py:50-72, 97, 124-142)."""
    def ns(name, sort_cols=1):
        _, rows = gu.load_golden(name)
        return rows

    rows = ns("toy_strand_ooe")
    got = {k: int(v["n"]) for k, v in rows.items()}
    assert got == {"('+', '+')": 1, "('+', '-')": 3, "('-', '+')": 1, "('-', '-')": 1, "all": 6}
    rows = ns("toy_strand_igo")
    assert {k: int(v["n"]) for k, v in rows.items()} == {"('+', '+')": 1, "('+', '-')": 4, "('-', '-')": 1, "all": 6}
    rows = ns("toy_strand_dist_ctrl")
    assert sorted(int(v["n"]) for v in rows.values()) == [1, 1, 1, 1, 2, 6]
    z, rows = gu.load_golden("toy_stripes")
    assert list(z["row0.vertical_stripe"][0]) == [22015, 11287, 10852, 11376, 9998]
    assert list(z["row0.coordinates"][0]) == ["chr1", "102000000", "102500000", "chr1", "105000000", "105500000"]


def test_legacy_loop_ref_is_statistically_consistent(fixtures_dir):
    """tests/loop_ref.np.txt is a legacy Monte-Carlo output (nshifts=10, seed 0, old CLI): it cannot be
    reproduced bit-for-bit, but the control-normalised loop pile-up must agree statistically."""
    ref = np.loadtxt(os.path.join(fixtures_dir, "loop_ref.np.txt"))
    _, rows = gu.load_golden("scc1_loops_ctrl")
    mine = rows["all"]["data"]
    assert ref.shape == mine.shape == (21, 21)
    m = np.isfinite(ref) & np.isfinite(mine)
    r = np.corrcoef(ref[m], mine[m])[0, 1]
    assert r > 0.9
    assert abs(mine[10, 10] / ref[10, 10] - 1) < 0.1


def test_legacy_loop_ref_with_its_own_options(fixtures_dir):
    """The options in loop_ref.np.txt's header (tests/loop_ref.np.txt:1-33: nshifts 10, seed 0, coverage_norm,
    unbalanced, mindist 0, pad 100 kb) through the restated current algorithm: coverage computed like
    cooltools.coverage (expected.coverage), 10 random-shift controls.  A pre-1.0 CLI with another control generator
    made the file, so agreement is statistical: Pearson 0.93, centre 1.62 vs 1.80, median |rel| 3 %."""
    import pandas as pd

    from coolpuppy_b200.expected import coverage

    clr = Cooler(os.path.join(fixtures_dir, "Scc1-control.10000.cool"))
    cis, tot = coverage(clr, 2)
    clr.add_bin_column("cov_cis_raw", cis)
    clr.add_bin_column("cov_tot_raw", tot)
    loops = pd.read_csv(os.path.join(fixtures_dir, "CH12_loops_Rao.bed"), sep="\t", header=None).iloc[:, :6]
    loops.columns = ["chrom1", "start1", "end1", "chrom2", "start2", "end2"]
    ref = np.loadtxt(os.path.join(fixtures_dir, "loop_ref.np.txt"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = oracle_pileup(clr, loops, features_format="bedpe", clr_weight_name=None, flank=100_000, mindist=0,
                            nshifts=10, seed=0, coverage_norm=True)
    mine = np.asarray(res.rows[0]["data"], dtype=float)
    assert np.corrcoef(mine.ravel(), ref.ravel())[0, 1] > 0.9
    assert abs(mine[10, 10] / ref[10, 10] - 1) < 0.15
    assert np.median(np.abs(mine - ref) / ref) < 0.06

"""expected_cis (numpy and GPU) pinned to the cooltools-made table the reference ships.

``tests/fixtures/CN.mm9.toy_expected.tsv`` is a copy of the reference's ``tests/data/CN.mm9.toy_expected.tsv``
(made by ``cooltools expected-cis`` on ``CN.mm9.1000kb.cool`` with the ``foo`` / ``bar`` view).  ``balanced.sum``
also pins the balanced pixel fetch (``w[row] * w[col] * count``) from outside this repository.
"""
import os

import numpy as np
import pandas as pd
import pytest

from coolpuppy_b200.coolio import Cooler

COLS = ["n_valid", "count.sum", "balanced.sum", "balanced.avg"]


def _inputs(fixtures_dir):
    clr = Cooler(os.path.join(fixtures_dir, "CN.mm9.1000kb.cool"))
    view = pd.read_csv(os.path.join(fixtures_dir, "CN.mm9.toy_regions.bed"), sep="\t", header=None,
                       names=["chrom", "start", "end", "name"])
    table = pd.read_csv(os.path.join(fixtures_dir, "CN.mm9.toy_expected.tsv"), sep="\t")
    return clr, view, table


def _check(ours, table):
    assert len(ours) == len(table) == 100
    assert list(ours["region1"]) == list(table["region1"]) and list(ours["dist"]) == list(table["dist"])
    assert np.array_equal(ours["n_valid"].values, table["n_valid"].values)
    for c in COLS[1:]:
        a, b = ours[c].values.astype(float), table[c].values.astype(float)
        assert np.array_equal(np.isnan(a), np.isnan(b)), c
        m = ~np.isnan(b)
        np.testing.assert_allclose(a[m], b[m], rtol=1e-12, err_msg=c)
    # count.avg is count.sum / n_valid (cooltools), NaN where the sum is masked
    with np.errstate(invalid="ignore", divide="ignore"):
        np.testing.assert_allclose(ours["count.avg"].values, ours["count.sum"].values / ours["n_valid"].values,
                                   rtol=0, equal_nan=True)


def test_expected_cis_matches_cooltools_table(fixtures_dir):
    from coolpuppy_b200.expected import expected_cis

    clr, view, table = _inputs(fixtures_dir)
    # region bar has NaN-weight bins: count.sum must still include their pixels (48 of its rows differ otherwise)
    assert np.isnan(np.asarray(clr.bins()["weight"].fetch(("chr2", 100_000_000, 150_000_000)).values)).any()
    _check(expected_cis(clr, view_df=view, clr_weight_name="weight", ignore_diags=2), table)


def test_raw_expected_counts_every_pixel(fixtures_dir):
    """Without weights every bin is valid: n_valid = nb - d and count.sum equals the balanced run's count.sum."""
    from coolpuppy_b200.expected import expected_cis

    clr, view, table = _inputs(fixtures_dir)
    raw = expected_cis(clr, view_df=view, clr_weight_name=None, ignore_diags=2)
    m = ~np.isnan(table["count.sum"].values)
    np.testing.assert_allclose(raw["count.sum"].values[m], table["count.sum"].values[m], rtol=0)
    assert np.array_equal(raw["n_valid"].values, np.concatenate([50 - np.arange(50), 50 - np.arange(50)]))


@pytest.mark.gpu
def test_expected_cis_gpu_matches_cooltools_table(fixtures_dir):
    from coolpuppy_b200.expected import expected_cis_gpu

    clr, view, table = _inputs(fixtures_dir)
    _check(expected_cis_gpu(clr, view_df=view, clr_weight_name="weight", ignore_diags=2), table)


def test_coverage_matches_cooltools_columns(fixtures_dir):
    """expected.coverage (what the reference gets from cooltools.coverage when coverage_norm is requested,
    coolpup.py:955-963) == the cov_cis_raw / cov_tot_raw columns cooltools stored in the reference's own fixture."""
    from coolpuppy_b200.expected import coverage

    clr = Cooler(os.path.join(fixtures_dir, "CN.mm9.1000kb.cool"))
    cis, tot = coverage(clr, ignore_diags=2)
    assert np.array_equal(cis, clr._bin_column("cov_cis_raw"))
    assert np.array_equal(tot, clr._bin_column("cov_tot_raw"))

"""Golden-vector case definitions shared by `make_golden.py` (runs the real reference, build container only)
and by the parity tests (run the oracle / the CUDA path on the same inputs, anywhere)."""
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(os.path.dirname(HERE), "fixtures")

BED6 = ["chrom", "start", "end", "name", "score", "strand"]
BEDPE6 = ["chrom1", "start1", "end1", "chrom2", "start2", "end2"]

TOY = dict(cooler="CN.mm9.1000kb.cool", features="toy_features.bed", features_schema="bed6", view="CN.mm9.toy_regions.bed")
TOYKW = dict(features_format="bed", flank=2_000_000, mindist=0)
SCC1 = dict(cooler="Scc1-control.10000.cool")

CASES = {
    # ---- toy fixture: the configurations of the reference's own tests (tests/test_coolpup.py:19-172)
    "toy_strand_ooe": {**TOY, "expected": "CN.mm9.toy_expected.tsv", "kwargs": {**TOYKW, "by_strand": True, "ooe": True}},
    "toy_strand_notooe": {**TOY, "expected": "CN.mm9.toy_expected.tsv", "kwargs": {**TOYKW, "by_strand": True, "ooe": False}},
    "toy_strand_balanced": {**TOY, "kwargs": {**TOYKW, "by_strand": True}},
    "toy_strand_rawcov": {**TOY, "kwargs": {**TOYKW, "by_strand": True, "clr_weight_name": None, "coverage_norm": True}},
    "toy_strand_igo": {**TOY, "kwargs": {**TOYKW, "by_strand": True, "ignore_group_order": True}},
    "toy_flipneg": {**TOY, "kwargs": {**TOYKW, "by_strand": True, "flip_negative_strand": True}},
    "toy_flipneg_igo": {**TOY, "kwargs": {**TOYKW, "by_strand": True, "flip_negative_strand": True, "ignore_group_order": True}},
    "toy_controls": {**TOY, "kwargs": {**TOYKW, "by_strand": True, "nshifts": 3, "seed": 0}},
    "toy_strand_dist_ctrl": {**TOY, "kwargs": {**TOYKW, "by_strand": True, "by_distance": True, "nshifts": 1, "seed": 1}},
    "toy_dist_edges": {**TOY, "by_distance_edges": [0, 2_000_000, 4_000_000, 8_000_000],
                       "kwargs": {**TOYKW, "nshifts": 2, "seed": 5}},
    "toy_bywindow": {**TOY, "kwargs": {**TOYKW, "by_window": True}},
    "toy_stripes": {**TOY, "kwargs": {**TOYKW, "store_stripes": True, "clr_weight_name": None, "min_diag": 0}},
    "toy_stripes_strand_ooe": {**TOY, "expected": "CN.mm9.toy_expected.tsv",
                               "kwargs": {**TOYKW, "by_strand": True, "ooe": True, "store_stripes": True}},
    "toy_stripes_notooe": {**TOY, "expected": "CN.mm9.toy_expected.tsv",
                           "kwargs": {**TOYKW, "by_distance": True, "ooe": False, "store_stripes": True}},
    "toy_stripes_ctrl_local": {**TOY, "kwargs": {"features_format": "bed", "flank": 3_000_000, "local": True, "nshifts": 2,
                                                  "seed": 9, "store_stripes": True}},
    "toy_wholechrom": {**{k: v for k, v in TOY.items() if k != "view"}, "kwargs": {"features_format": "bed", "flank": 3_000_000, "mindist": 0}},
    "toy_local_ooe": {**TOY, "expected": "CN.mm9.toy_expected.tsv", "kwargs": {"features_format": "bed", "flank": 2_000_000, "local": True, "ooe": True}},
    "toy_local_raw": {**TOY, "kwargs": {"features_format": "bed", "flank": 3_000_000, "local": True, "clr_weight_name": None, "nshifts": 2, "seed": 2}},
    "toy_mindist_auto": {**TOY, "kwargs": {"features_format": "bed", "flank": 1_000_000}},
    # ---- Scc1 10 kb fixture (BASELINE.json configs[0], configs[1])
    "scc1_loops_raw": {**SCC1, "features": "CH12_loops_Rao.bed", "features_schema": "bedpe6",
                       "kwargs": dict(features_format="bedpe", clr_weight_name=None, flank=100_000, mindist=0, nshifts=0)},
    "scc1_loops_auto": {**SCC1, "features": "CH12_loops_Rao.bed", "features_schema": "bedpe6",
                        "kwargs": dict(features_format="bedpe", clr_weight_name=None, flank=100_000)},
    "scc1_loops_pad41": {**SCC1, "features": "CH12_loops_Rao.bed", "features_schema": "bedpe6",
                         "kwargs": dict(features_format="bedpe", clr_weight_name=None, flank=410_000, mindist=0)},
    "scc1_loops_ctrl": {**SCC1, "features": "CH12_loops_Rao.bed", "features_schema": "bedpe6",
                        "kwargs": dict(features_format="bedpe", clr_weight_name=None, flank=100_000, mindist=0, nshifts=10, seed=0)},
    "scc1_loops_dist": {**SCC1, "features": "CH12_loops_Rao.bed", "features_schema": "bedpe6",
                        "kwargs": dict(features_format="bedpe", clr_weight_name=None, flank=100_000, by_distance=True, nshifts=1, seed=7)},
    "scc1_ctcf_local_ooe": {**SCC1, "features": "Bonev_CTCF+.bed", "features_schema": "bed3", "expected": "compute:count.avg",
                            "kwargs": dict(features_format="bed", clr_weight_name=None, flank=100_000, local=True,
                                           expected_value_col="count.avg", ooe=True)},
    "scc1_ctcf_local_subset_notooe": {**SCC1, "features": "Bonev_CTCF+.bed", "features_schema": "bed3", "expected": "compute:count.avg",
                                      "kwargs": dict(features_format="bed", clr_weight_name=None, flank=200_000, local=True,
                                                     expected_value_col="count.avg", ooe=False, subset=4000, seed=11)},
    "scc1_ctcf_pairs_strand_dist": {**SCC1, "features": "ctcf_stranded_chr18_19.bed", "features_schema": "bed6",
                                    "kwargs": dict(features_format="bed", clr_weight_name=None, flank=100_000, by_strand=True,
                                                   by_distance=True, nshifts=2, seed=3)},
    "scc1_ctcf_pairs_flip_ooe": {**SCC1, "features": "ctcf_stranded_chr18_19.bed", "features_schema": "bed6", "expected": "compute:count.avg",
                                 "kwargs": dict(features_format="bed", clr_weight_name=None, flank=100_000, by_strand=True,
                                                flip_negative_strand=True, expected_value_col="count.avg", ooe=True, maxdist=5_000_000)},
    "scc1_loops_stripes_ctrl": {**SCC1, "features": "CH12_loops_Rao.bed", "features_schema": "bedpe6",
                                "kwargs": dict(features_format="bedpe", clr_weight_name=None, flank=50_000, nshifts=1, seed=5,
                                               maxdist=400_000, store_stripes=True)},
    # ---- ignore_diags other than the default 2 (0: only the lower triangle is masked, the diagonal stays)
    "toy_igd0_strand": {**TOY, "kwargs": {**TOYKW, "by_strand": True, "min_diag": 0}},
    "toy_igd3_ooe_stripes": {**TOY, "expected": "CN.mm9.toy_expected.tsv",
                             "kwargs": {**TOYKW, "ooe": True, "min_diag": 3, "store_stripes": True}},
    "toy_local_igd0_raw_ctrl": {**TOY, "kwargs": {"features_format": "bed", "flank": 3_000_000, "local": True, "clr_weight_name": None,
                                                 "min_diag": 0, "nshifts": 2, "seed": 6}},
    "scc1_loops_igd0": {**SCC1, "features": "CH12_loops_Rao.bed", "features_schema": "bedpe6",
                        "kwargs": dict(features_format="bedpe", clr_weight_name=None, flank=100_000, mindist=0, min_diag=0)},
    "scc1_loops_igd5_ctrl": {**SCC1, "features": "CH12_loops_Rao.bed", "features_schema": "bedpe6",
                             "kwargs": dict(features_format="bedpe", clr_weight_name=None, flank=100_000, mindist=0, min_diag=5,
                                            nshifts=2, seed=8)},
    # ---- expected == 0 under non-zero counts (x / 0 = +inf): the reference's merge of >= 2 regions / groups turns
    # +inf into 1.797e308 (np.nan_to_num in sum_pups, lib/puputils.py:97-98) instead of NaN
    "toy_zero_expected_ooe": {**TOY, "expected": "CN.mm9.toy_expected.tsv", "expected_zero": {"foo": [3, 4], "bar": [4]},
                              "kwargs": {**TOYKW, "ooe": True}},
    "toy_zero_expected_strand": {**TOY, "expected": "CN.mm9.toy_expected.tsv", "expected_zero": {"foo": [3, 4], "bar": [4, 6]},
                                 "kwargs": {**TOYKW, "ooe": True, "by_strand": True}},
    "toy_zero_expected_bywindow": {**TOY, "expected": "CN.mm9.toy_expected.tsv", "expected_zero": {"foo": [4, 5], "bar": [3, 4]},
                                   "kwargs": {**TOYKW, "ooe": True, "by_window": True}},
    "toy_zero_expected_flip_dist": {**TOY, "expected": "CN.mm9.toy_expected.tsv", "expected_zero": {"bar": [2, 3, 4, 5, 6, 7]},
                                    "kwargs": {**TOYKW, "ooe": True, "by_strand": True, "by_distance": True,
                                               "flip_negative_strand": True}},
    "toy_zero_expected_one_region": {**TOY, "expected": "CN.mm9.toy_expected.tsv", "expected_zero": {"foo": [3]},
                                     "view_rows": [0], "kwargs": {**TOYKW, "ooe": True}},
    # ---- trans (inter-chromosomal) pile-ups: rectangular region1 x region2 matrices, scalar expected, no diagonal
    # mask (coolpup.py:652-680, 999-1005, 1126-1128, 1419-1426)
    "toy_trans_raw": {**TOY, "kwargs": {"features_format": "bed", "flank": 2_000_000, "trans": True, "clr_weight_name": None}},
    "toy_trans_strand": {**TOY, "kwargs": {"features_format": "bed", "flank": 2_000_000, "trans": True, "by_strand": True}},
    "toy_trans_ctrl": {**TOY, "kwargs": {"features_format": "bed", "flank": 2_000_000, "trans": True, "nshifts": 2, "seed": 4}},
    "toy_trans_expected_ooe": {**TOY, "expected": "toy_trans_expected.tsv",
                               "kwargs": {"features_format": "bed", "flank": 3_000_000, "trans": True, "ooe": True}},
    "toy_trans_notooe_flip": {**TOY, "expected": "toy_trans_expected.tsv",
                              "kwargs": {"features_format": "bed", "flank": 2_000_000, "trans": True, "ooe": False,
                                         "by_strand": True, "flip_negative_strand": True}},
    "toy_trans_bywindow_rawcov": {**{k: v for k, v in TOY.items() if k != "view"},
                                  "kwargs": {"features_format": "bed", "flank": 2_000_000, "trans": True, "by_window": True,
                                             "clr_weight_name": None, "coverage_norm": True}},
    "toy_trans_bedpe_ctrl": {**TOY, "features": "toy_trans.bedpe", "features_schema": "bedpe6",
                             "kwargs": {"features_format": "bedpe", "flank": 1_000_000, "trans": True, "nshifts": 2, "seed": 12}},
    # ---- rescaled pile-ups: windows of the features' own sizes (+ rescale_flank of it on either side) zoomed to
    # rescale_size x rescale_size (coolpup.py:1193-1234, 87-90, 108-114); "toy_sized_features.bed" = toy features 2-5 Mb long
    "toy_rescale_local": {**TOY, "features": "toy_sized_features.bed",
                          "kwargs": {"features_format": "bed", "local": True, "rescale": True, "rescale_flank": 1, "rescale_size": 9}},
    "toy_rescale_pairs_strand": {**TOY, "features": "toy_sized_features.bed",
                                 "kwargs": {"features_format": "bed", "mindist": 0, "rescale": True, "rescale_flank": 0.5,
                                            "rescale_size": 9, "by_strand": True}},
    "toy_rescale_ooe": {**TOY, "features": "toy_sized_features.bed", "expected": "CN.mm9.toy_expected.tsv",
                        "kwargs": {"features_format": "bed", "mindist": 0, "rescale": True, "rescale_flank": 1, "rescale_size": 11,
                                   "ooe": True}},
    "toy_rescale_notooe_local": {**TOY, "features": "toy_sized_features.bed", "expected": "CN.mm9.toy_expected.tsv",
                                 "kwargs": {"features_format": "bed", "local": True, "rescale": True, "rescale_flank": 1,
                                            "rescale_size": 9, "ooe": False}},
    "toy_rescale_cov_ctrl": {**TOY, "features": "toy_sized_features.bed",
                             "kwargs": {"features_format": "bed", "mindist": 0, "rescale": True, "rescale_flank": 0.3,
                                        "rescale_size": 7, "clr_weight_name": None, "coverage_norm": True, "nshifts": 2, "seed": 3}},
    "toy_rescale_up": {**TOY, "kwargs": {"features_format": "bed", "mindist": 0, "rescale": True, "rescale_flank": 1, "rescale_size": 15}},
    "toy_rescale_bedpe_ctrl": {**TOY, "features": "toy_sized.bedpe", "features_schema": "bedpe6",
                               "kwargs": {"features_format": "bedpe", "mindist": 0, "rescale": True, "rescale_flank": 0.5,
                                          "rescale_size": 9, "nshifts": 2, "seed": 4}},
    "toy_rescale_bywindow": {**TOY, "features": "toy_sized_features.bed",
                             "kwargs": {"features_format": "bed", "mindist": 0, "rescale": True, "rescale_flank": 1,
                                        "rescale_size": 9, "by_window": True}},
    "toy_rescale_flip_dist": {**TOY, "features": "toy_sized_features.bed", "expected": "CN.mm9.toy_expected.tsv",
                              "kwargs": {"features_format": "bed", "mindist": 0, "rescale": True, "rescale_flank": 0.5,
                                         "rescale_size": 11, "by_strand": True, "flip_negative_strand": True, "ooe": True}},
    "scc1_tads_rescale_ctrl": {**SCC1, "features": "CH12_TADs_Rao.bed", "features_schema": "bed3",
                               "kwargs": dict(features_format="bed", clr_weight_name=None, local=True, rescale=True, rescale_flank=1,
                                              rescale_size=33, subset=150, seed=2, nshifts=3)},
    "scc1_tads_rescale_local": {**SCC1, "features": "CH12_TADs_Rao.bed", "features_schema": "bed3",
                                "kwargs": dict(features_format="bed", clr_weight_name=None, local=True, rescale=True, rescale_flank=1,
                                               rescale_size=99, subset=400, seed=1)},
    "scc1_ctcf_pairs_arms": {**SCC1, "features": "ctcf_stranded_chr18_19.bed", "features_schema": "bed6", "view": "scc1_arms_view.bed",
                             "kwargs": dict(features_format="bed", clr_weight_name=None, flank=50_000, mindist=0, maxdist=2_000_000,
                                            nshifts=1, seed=4)},
}


def load_features(spec):
    path = os.path.join(FIX, spec["features"])
    df = pd.read_csv(path, sep="\t", header=None, comment="#")
    schema = spec.get("features_schema", "bed3")
    names = {"bed3": BED6[:3], "bed6": BED6, "bedpe6": BEDPE6}[schema]
    df = df.iloc[:, : len(names)]
    df.columns = names
    return df


def load_view(spec):
    if "view" not in spec:
        return None
    df = pd.read_csv(os.path.join(FIX, spec["view"]), sep="\t", header=None)
    df.columns = ["chrom", "start", "end", "name"]
    if "view_rows" in spec:
        df = df.iloc[spec["view_rows"]].reset_index(drop=True)
    return df


def load_expected(spec, clr, view, expected_cis_func):
    e = spec.get("expected")
    if e is None:
        return None
    if e.startswith("compute:"):
        return expected_cis_func(clr, view_df=view, clr_weight_name=None, ignore_diags=2)
    tab = pd.read_csv(os.path.join(FIX, e), sep="\t", dtype={"region1": str, "region2": str})
    for region, dists in spec.get("expected_zero", {}).items():  # zero the expected at these distances
        hit = (tab["region1"] == region) & (tab["region2"] == region) & tab["dist"].isin(dists)
        for c in tab.columns:
            if c.endswith(".avg") or c.endswith(".sum"):
                tab.loc[hit, c] = 0.0
    return tab


def write_derived_fixtures():
    """Derived feature tables (subsets of the reference's CTCF lists) used by the pair-wise cases."""
    plus = pd.read_csv(os.path.join(FIX, "Bonev_CTCF+.bed"), sep="\t", header=None).iloc[:, :3]
    minus = pd.read_csv(os.path.join(FIX, "Bonev_CTCF-.bed"), sep="\t", header=None).iloc[:, :3]
    plus.columns = minus.columns = ["chrom", "start", "end"]
    rows = []
    rng = np.random.default_rng(20261017)
    for chrom, n in (("chr18", 90), ("chr19", 70)):
        for df, strand in ((plus, "+"), (minus, "-")):
            sub = df[df.chrom == chrom]
            idx = np.sort(rng.choice(len(sub), size=n, replace=False))
            sub = sub.iloc[idx].copy()
            sub["name"] = "ctcf"
            sub["score"] = 0
            sub["strand"] = strand
            rows.append(sub)
    out = pd.concat(rows).sort_values(["chrom", "start"]).reset_index(drop=True)
    out.to_csv(os.path.join(FIX, "ctcf_stranded_chr18_19.bed"), sep="\t", header=False, index=False)
    arms = pd.DataFrame(
        [
            ("chr18", 0, 40_000_000, "chr18_p"),
            ("chr18", 40_000_000, 90_772_031, "chr18_q"),
            ("chr19", 3_000_000, 61_342_430, "chr19_main"),
        ]
    )
    arms.to_csv(os.path.join(FIX, "scc1_arms_view.bed"), sep="\t", header=False, index=False)


if __name__ == "__main__":
    write_derived_fixtures()

#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (`/root/reference/coolpuppy`).

Only runnable in the build container (needs `/root/reference`).  The reference's absent third-party imports
(cooler, cooltools, bioframe, natsort, more_itertools) are satisfied by `oracle/refshim/` (see its README).  The
reference source itself is imported as it lies under `/root/reference` -- nothing is copied.

Outputs (committed): `tests/golden/<case>.npz` + `tests/golden/manifest.json`.  For every case we store

* the final DataFrame rows: group key (repr), `data`, `num`, `n`, `control_n`, `control_num`
  (and stripes / coordinates when `store_stripes`),
* for every view region: the window stream exactly as the reference's `CoordCreator.pos_stream` emitted it
  (chromosome-relative `stBin1`, `stBin2`, kind, group key, flip flag, in emission order -- this pins pair
  enumeration order, distance filtering and the `np.random` draw order of the control shifts), and
* the raw per-region accumulators returned by `PileUpper.pileup_region` (`data` sum, `num`, `n`, `cov_start`,
  `cov_end` per kind and group) -- the boundary the C-ABI replaces (`coolpup.py:1285-1358`).

Usage:  python tests/golden/make_golden.py [case ...]
"""
import json
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
FIX = os.path.join(REPO, "tests", "fixtures")
sys.path.insert(0, os.path.join(REPO, "oracle", "refshim"))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

warnings.simplefilter("ignore")

import cooler  # noqa: E402  (refshim)
from coolpuppy import coolpup as ref  # noqa: E402  (the real reference)
from coolpuppy_b200.expected import expected_cis  # noqa: E402
from golden_cases import CASES, load_features, load_view, load_expected  # noqa: E402

# ---------------------------------------------------------------------------------------------------------------
# recording wrappers around the reference's own generator / per-region function
REC = {"region": None, "windows": {}, "regions": {}}


def _key(g):
    if isinstance(g, str):
        return g
    return repr(tuple(x.item() if isinstance(x, np.generic) else x for x in g))


def _wrap_stream(orig):
    def rec(self, *a, **k):
        for row in orig(self, *a, **k):
            if row is not None:
                REC["windows"].setdefault(REC["region"], []).append(
                    (
                        int(row["stBin1"]),
                        int(row["stBin2"]),
                        0 if row["kind"] == "ROI" else 1,
                        _key(row["group"]),
                        int(bool(row.get("flip", False))),
                    )
                )
            yield row

    return rec


ref.CoordCreator.get_combinations = _wrap_stream(ref.CoordCreator.get_combinations)
ref.CoordCreator.get_intervals_stream = _wrap_stream(ref.CoordCreator.get_intervals_stream)
_orig_region = ref.PileUpper.pileup_region


def _rec_region(self, region1, region2=None, *a, **k):
    REC["region"] = region1
    out = _orig_region(self, region1, region2, *a, **k)
    snap = {}
    for kind in ("ROI", "control"):
        for g, pup in out[kind].items():
            snap[(kind, _key(g))] = {
                "data": np.array(pup["data"], dtype=float),
                "num": np.array(pup["num"]),
                "n": int(pup["n"]),
                "cov_start": np.array(pup["cov_start"], dtype=float),
                "cov_end": np.array(pup["cov_end"], dtype=float),
            }
    REC["regions"][region1] = snap
    return out


ref.PileUpper.pileup_region = _rec_region


def run_case(name, spec):
    REC["windows"].clear()
    REC["regions"].clear()
    clr = cooler.Cooler(os.path.join(FIX, spec["cooler"]))
    features = load_features(spec)
    kw = dict(spec["kwargs"])
    if "by_distance_edges" in spec:
        kw["by_distance"] = np.asarray(spec["by_distance_edges"])
    view = load_view(spec)
    if view is not None:
        kw["view_df"] = view
    exp = load_expected(spec, clr, view, expected_cis)
    if exp is not None:
        kw["expected_df"] = exp
    pups = ref.pileup(clr, features, **kw)

    out = {}
    keys = []
    sort_cols = [c for c in ("orientation", "distance_band", "chrom", "start", "end", "group") if c in pups.columns]
    for i, row in pups.iterrows():
        if "group" in pups.columns:
            k = _key(row["group"])
        else:  # by-window output has no group column
            k = repr((row["chrom"], int(row["start"]), int(row["end"])))
        keys.append(k)
        out[f"row{i}.data"] = np.asarray(row["data"], dtype=float)
        out[f"row{i}.num"] = np.asarray(row["num"])
        out[f"row{i}.n"] = np.int64(row["n"])
        if "control_n" in pups.columns:
            out[f"row{i}.control_n"] = np.int64(row["control_n"])
            out[f"row{i}.control_num"] = np.asarray(row["control_num"])
        if spec["kwargs"].get("store_stripes"):
            out[f"row{i}.vertical_stripe"] = np.asarray(row["vertical_stripe"], dtype=float)
            out[f"row{i}.horizontal_stripe"] = np.asarray(row["horizontal_stripe"], dtype=float)
            out[f"row{i}.coordinates"] = np.asarray(row["coordinates"]).astype(str)
    out["row_keys"] = np.asarray(keys)
    for extra in ("orientation", "separation"):
        if extra in pups.columns:
            out[f"col.{extra}"] = np.asarray(pups[extra].astype(str))
    # scalar annotation columns a consumer (plotpup / save_pileup_df) reads
    ann = {}
    for c in pups.columns:
        v = pups[c].iloc[0]
        if isinstance(v, (str, bool, int, float, np.integer, np.floating, np.bool_)) and c not in ("n",):
            ann[c] = v.item() if isinstance(v, np.generic) else v
    out["columns"] = np.asarray(list(pups.columns))
    out["annotations_json"] = np.asarray(json.dumps(ann, default=str))

    regions = list(REC["regions"].keys())
    out["regions"] = np.asarray(regions)
    for r in regions:
        w = REC["windows"].get(r, [])
        out[f"win.{r}.st1"] = np.asarray([x[0] for x in w], dtype=np.int64)
        out[f"win.{r}.st2"] = np.asarray([x[1] for x in w], dtype=np.int64)
        out[f"win.{r}.kind"] = np.asarray([x[2] for x in w], dtype=np.int8)
        out[f"win.{r}.flip"] = np.asarray([x[4] for x in w], dtype=np.int8)
        gk = sorted({x[3] for x in w})
        gid = {g: i for i, g in enumerate(gk)}
        out[f"win.{r}.group_keys"] = np.asarray(gk) if gk else np.asarray([], dtype=str)
        out[f"win.{r}.group"] = np.asarray([gid[x[3]] for x in w], dtype=np.int32)
        snap = REC["regions"][r]
        out[f"acc.{r}.keys"] = np.asarray([f"{k[0]}|{k[1]}" for k in snap]) if snap else np.asarray([], dtype=str)
        for j, (k, pup) in enumerate(snap.items()):
            for f in ("data", "num", "cov_start", "cov_end"):
                out[f"acc.{r}.{j}.{f}"] = pup[f]
            out[f"acc.{r}.{j}.n"] = np.int64(pup["n"])
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    nwin = sum(len(v) for v in REC["windows"].values())
    print(f"{name}: rows={len(keys)} windows={nwin} n_all={[int(out[f'row{i}.n']) for i in range(len(keys))][-1]}")
    return {"rows": len(keys), "windows": nwin}


def main():
    which = sys.argv[1:] or list(CASES)
    manifest_path = os.path.join(HERE, "manifest.json")
    manifest = {}
    if os.path.exists(manifest_path):
        manifest = json.load(open(manifest_path))
    for name in which:
        spec = CASES[name]
        info = run_case(name, spec)
        manifest[name] = {"spec": spec, **info}
    json.dump(manifest, open(manifest_path, "w"), indent=1, sort_keys=True)
    versions = {"numpy": np.__version__, "pandas": pd.__version__, "reference": "open2c/coolpuppy 1.1.0 @592673c"}
    json.dump(versions, open(os.path.join(HERE, "versions.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

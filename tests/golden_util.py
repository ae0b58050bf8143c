"""Helpers shared by the parity tests: load a golden case and rebuild its inputs."""
import json
import os

import numpy as np

from golden_cases import CASES, FIX, load_expected, load_features, load_view

from coolpuppy_b200.coolio import Cooler
from coolpuppy_b200.expected import expected_cis

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_coolers = {}


def get_cooler(name):
    if name not in _coolers:
        _coolers[name] = Cooler(os.path.join(FIX, name))
    return _coolers[name]


def case_inputs(name):
    """(clr, features, kwargs) exactly as make_golden.py passed them to the reference's pileup()."""
    spec = CASES[name]
    clr = get_cooler(spec["cooler"])
    kw = dict(spec["kwargs"])
    if "by_distance_edges" in spec:
        kw["by_distance"] = np.asarray(spec["by_distance_edges"])
    view = load_view(spec)
    if view is not None:
        kw["view_df"] = view
    exp = load_expected(spec, clr, view, expected_cis)
    if exp is not None:
        kw["expected_df"] = exp
    return clr, load_features(spec), kw


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False)
    rows = {}
    for i, k in enumerate(z["row_keys"]):
        r = {f.split(".", 1)[1]: z[f] for f in z.files if f.startswith(f"row{i}.")}
        rows[str(k)] = r
    return z, rows


def all_cases():
    return list(CASES)


def manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))

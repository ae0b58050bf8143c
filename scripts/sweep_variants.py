"""Tuning sweep of the main pile-up kernel on the bench workload (run on the GPU box).

One process: the synthetic genome and the window lists are generated once; for every variant
``R:S[:bucket_target[:chunk[:prefetch]]][;KEY=VALUE...]`` (strip height, lanes per strip run, pixels per bucket,
windows per chunk; ``KEY=VALUE`` sets the tuning variable ``PUP_KEY``, e.g. ``SCHED=0``, ``MINB=3``) the regions are
re-indexed (the strip geometry is fixed at region creation), the full pass is timed with CUDA events and the
accumulators are compared with the first variant's: ``num`` / ``n`` exactly, ``sum`` to 1e-9 relative.

    python scripts/sweep_variants.py 1:4 2:8 2:16 4:16 8:32 --chroms all --steps 3
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import bench
from coolpuppy_b200 import _native
from coolpuppy_b200.synthetic import synthetic_region


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="+")
    ap.add_argument("--chroms", default="all")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--depth", type=float, default=500.0)
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--nshifts", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    _native.require_device()
    sizes = bench.chromsizes(a)
    names = list(sizes)
    a.workload = "configs3"
    features, n_pairs = bench.make_features(a, sizes)
    windows, _, _, _ = bench.build_windows(a, sizes, features, None)
    stream = torch.cuda.current_stream(dev).cuda_stream
    data, dwin = {}, {}
    for ci, c in enumerate(names):
        t = synthetic_region(windows[c]["nb"], depth=a.depth, seed=1234 + ci, device=dev, nan_frac=0.03)
        data[c] = {k: t[k] for k in ("nb", "indptr", "col", "count", "weight")}
        dwin[c] = tuple(torch.from_numpy(windows[c][k]).to(dev) for k in ("r0", "c0", "slot"))
        del t
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    W, n_slots = 2 * (bench.WORKLOADS["configs3"]["flank"] // bench.BINSIZE) + 1, 2
    stride = _native.acc_stride(W)
    ref = None
    rows = []
    for v in a.variants:
        toks = v.split(";")
        parts = toks[0].split(":")
        for k in [k for k in os.environ if k.startswith("PUP_") and not k.startswith("PUP_BENCH")]:
            del os.environ[k]  # every variant starts from the defaults
        os.environ["PUP_STRIP"], os.environ["PUP_LANES"] = parts[0], parts[1]
        if len(parts) > 2:
            os.environ["PUP_BUCKET_TARGET"] = parts[2]
        if len(parts) > 3:
            os.environ["PUP_CHUNK"] = parts[3]
        if len(parts) > 4:
            os.environ["PUP_PREFETCH"] = parts[4]
        for t in toks[1:]:  # KEY=VALUE -> PUP_KEY
            k, val = t.split("=")
            os.environ["PUP_" + k] = val
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        regions = {c: _native.Region(0, d["nb"], d["indptr"], d["col"], d["count"], d["weight"], None, None, ignore_diags=2,
                                     flags=0, stream=stream) for c, d in data.items()}
        torch.cuda.synchronize()
        prep_ms = (time.perf_counter() - t0) * 1e3
        acc = torch.zeros(n_slots * stride, dtype=torch.float64, device=dev)

        def step():
            acc.zero_()
            for c in names:
                r0, c0, sl = dwin[c]
                regions[c].accumulate(r0, c0, sl, W, n_slots, 0, acc, stream=stream)

        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        _native.timing_enable(True)
        _native.timing_read(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        ph = _native.timing_read(reset=True)
        _native.timing_enable(False)
        out = _native.acc_export(acc, W, n_slots, device=0, stream=stream)
        nbytes = sum(r.device_bytes for r in regions.values())
        for r in regions.values():
            r.close()
        del regions
        torch.cuda.synchronize()
        ok = None
        if ref is None:
            ref = out
        else:
            m = np.isfinite(ref["sum"]) & (ref["sum"] != 0)
            rel = float(np.max(np.abs(out["sum"][m] - ref["sum"][m]) / np.abs(ref["sum"][m]))) if m.any() else 0.0
            ok = bool(np.array_equal(out["num"], ref["num"]) and np.array_equal(out["n"], ref["n"]) and rel < 1e-9
                      and np.array_equal(np.isfinite(out["sum"]), np.isfinite(ref["sum"])))
            ok = (ok, rel)
        row = {"variant": v, "ms_per_step": ms, "main_ms": (ph["main"][0] + ph["dense_band"][0]) / a.steps, "dense_ms": ph["dense_band"][0] / a.steps, "plan_ms": ph["plan"][0] / a.steps,
               "counts_ms": ph["vector"][0] / a.steps, "prep_ms": prep_ms, "region_gb": nbytes / 1e9,
               "n": int(out["n"].sum()), "matches_first": ok}
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

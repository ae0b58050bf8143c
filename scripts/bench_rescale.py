"""Rescaled local pile-up of TAD-like features on the synthetic 3 Gbp genome (second half):
wall time of ``pileup(clr, tads, local=True, rescale=True, rescale_flank=1, rescale_size=99)`` from host memory,
CUDA-event time of ``k_rescale``, and the restated reference path (oracle ``rescale_snip`` with the real scipy zoom) on
one host core over a sample of the same windows.  Run on the GPU box:

    python scripts/bench_rescale.py --tads 3000 --out gpurun_out/r2_bench_rescale.json
"""
import argparse
import json
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import pandas as pd
import torch

import bench
from coolpuppy_b200 import _native
from coolpuppy_b200 import coolpup as cp


def tads(n, sizes, seed=77, lo=200_000, hi=2_000_000):
    rng = np.random.default_rng(seed)
    names = list(sizes)
    L = np.array([sizes[c] for c in names], dtype=np.float64)
    ch = rng.choice(len(names), n, p=L / L.sum())
    length = (np.exp(rng.uniform(np.log(lo), np.log(hi), n)) // bench.BINSIZE * bench.BINSIZE).astype(np.int64)
    start = ((rng.random(n) * (L[ch] - 3 * length - 4 * bench.BINSIZE) + length + bench.BINSIZE) // bench.BINSIZE
             * bench.BINSIZE).astype(np.int64)
    df = pd.DataFrame({"chrom": [names[i] for i in ch], "start": start, "end": start + length})
    return df.sort_values(["chrom", "start"]).reset_index(drop=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tads", type=int, default=3000)
    ap.add_argument("--chroms", default="all")
    ap.add_argument("--depth", type=float, default=500.0)
    ap.add_argument("--rescale-size", type=int, default=99)
    ap.add_argument("--calls", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=150)
    ap.add_argument("--out", default="gpurun_out/r2_bench_rescale.json")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    _native.require_device()
    sizes = bench.chromsizes(a)
    _, host = bench.generate_genome(a, sizes, dev, keep_device=set(), keep_host=set(sizes), pin=True)
    clr = bench.host_cooler(sizes, host, pin=False)
    feats = tads(a.tads, sizes)
    kw = dict(features_format="bed", local=True, rescale=True, rescale_flank=1, rescale_size=a.rescale_size,
              clr_weight_name="weight", device=0)
    _native.timing_enable(True)
    times, kern = [], []
    for i in range(a.calls + 1):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pups = cp.pileup(clr, feats, **kw)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tm = _native.timing_read(reset=True)
        if i > 0:
            times.append(dt)
            kern.append(float(tm["main"][0]))
    n = int(pups["n"].iloc[0])
    sec = float(np.median(times))
    # CPU: the reference's per-snippet path on a sample of the windows of the largest chromosome
    from oracle.pileup_oracle import oracle_accumulate_rescaled

    c = max(sizes, key=sizes.get)
    h = host[c]
    ip = h["upper_indptr"].numpy().astype(np.int64)
    col = h["upper_col"].numpy()
    cnt_ = h["upper_count"].numpy()
    nb = h["nb"]
    from scipy import sparse

    up = sparse.csr_matrix((cnt_, col, ip), shape=(nb, int(col.max()) + 1))[:, :nb]
    full = (up + sparse.triu(up, 1).T).tocsr()
    f = feats[feats["chrom"] == c]
    sel = f.sample(min(a.cpu_sample, len(f)), random_state=1)
    length = (sel["end"] - sel["start"]).values
    r0 = ((sel["start"].values - length) // bench.BINSIZE).astype(np.int32)
    hh = (3 * length // bench.BINSIZE).astype(np.int32)
    z = np.zeros(0, dtype=np.int32)
    t0 = time.perf_counter()  # set-up cost of the restatement (CSR construction, balancing of the chromosome): untimed
    oracle_accumulate_rescaled(nb, full.indptr, full.indices, full.data, h["weight"].numpy(), None, None, z, z, z, z, z, None,
                               a.rescale_size, 2, 1, local=True)
    setup_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = oracle_accumulate_rescaled(nb, full.indptr, full.indices, full.data, h["weight"].numpy(), None, None, r0, r0, hh, hh,
                                     np.zeros(len(r0), dtype=np.int32), None, a.rescale_size, 2, 1, local=True)
    cpu_s = max(1e-9, time.perf_counter() - t0 - setup_s)
    cells = float(np.sum((3.0 * (feats["end"] - feats["start"]).values / bench.BINSIZE) ** 2))
    line = {
        "metric": "rescaled pile-ups/sec (TAD-like features, local, rescale_flank=1)", "unit": "pile-ups/s",
        "workload": f"{len(feats)} TAD-like features (200 kb - 2 Mb, log-uniform) on the synthetic 3 Gbp genome @10 kb, "
                    f"local=True, rescale_flank=1 (windows 60 - 600 bins), rescale_size={a.rescale_size}, balanced",
        "e2e_api": {"value": n / sec, "seconds_per_call": sec, "calls": len(times), "n": n,
                    "what": "pileup() from host memory: uploads + indexing of every chromosome, k_rescale, export"},
        "k_rescale": {"ms_per_call": float(np.median(kern)), "value": n / (float(np.median(kern)) / 1e3),
                      "dense_cells_per_call": cells, "cells_per_s": cells / (float(np.median(kern)) / 1e3)},
        "cpu_baseline": {"value": int(ref["n"].sum()) / cpu_s, "unit": "pile-ups/s", "cores": 1, "kind": "port",
                         "sample": f"{int(ref['n'].sum())} windows of {c} ({cpu_s:.1f} s), oracle _stream_snips + _rescale_snip "
                                   "(scipy.ndimage.zoom) + _add_snip; matrix fetch untimed"},
    }
    print(json.dumps(line))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(line, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

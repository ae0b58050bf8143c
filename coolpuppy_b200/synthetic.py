"""Synthetic Hi-C genomes for benchmarks and large-size parity properties.

There is no network, hence no real deep cooler: the benchmark genome is generated on the GPU with torch
(plumbing, not the product): hg38 chromosome lengths at 10 kb bins, upper-triangle counts
``~ Poisson(depth / s)`` at bin separation ``s >= 1`` and ``Poisson(depth)`` on the diagonal, mirrored into a
symmetric CSR; balancing weights ``exp(N(0, 0.2^2)) * 1e-2`` with a fraction of NaN bins; expected = per-diagonal
mean of the balanced values over valid bin pairs with the first two diagonals NaN (what ``cooltools expected-cis``
would produce).  Everything is seeded and deterministic for a given torch / GPU generation.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

HG38 = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555, "chr5": 181538259,
    "chr6": 170805979, "chr7": 159345973, "chr8": 145138636, "chr9": 138394717, "chr10": 133797422,
    "chr11": 135086622, "chr12": 133275309, "chr13": 114364328, "chr14": 107043718, "chr15": 101991189,
    "chr16": 90338345, "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415,
}


def synthetic_region(nb, depth=500.0, seed=0, device="cuda", nan_frac=0.03, rows_per_chunk=None):
    """One square region as device tensors: indptr/col/count (int32, symmetric CSR), weight, expected (f64)."""
    import torch

    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(int(seed))
    if rows_per_chunk is None:
        rows_per_chunk = max(1, min(nb, (32 << 20) // max(nb, 1)))
    ii, jj, cc = [], [], []
    cols = torch.arange(nb, device=dev, dtype=torch.int64)
    for a in range(0, nb, rows_per_chunk):
        b = min(nb, a + rows_per_chunk)
        rows = torch.arange(a, b, device=dev, dtype=torch.int64)
        sep = cols[None, :] - rows[:, None]
        rate = torch.where(sep >= 0, depth / sep.clamp(min=1).to(torch.float32), torch.zeros((), device=dev))
        cnt = torch.poisson(rate, generator=g)
        nz = cnt.nonzero(as_tuple=False)
        ii.append(nz[:, 0] + a)
        jj.append(nz[:, 1])
        cc.append(cnt[nz[:, 0], nz[:, 1]].to(torch.int32))
        del sep, rate, cnt, nz
    i = torch.cat(ii)
    j = torch.cat(jj)
    c = torch.cat(cc)
    del ii, jj, cc
    # weights and expected from the upper triangle
    w = torch.exp(torch.randn(nb, generator=g, device=dev, dtype=torch.float64) * 0.2) * 1e-2
    if nan_frac > 0:
        w[torch.rand(nb, generator=g, device=dev) < nan_frac] = float("nan")
    ok = ~torch.isnan(w)
    val = w[i] * w[j] * c.to(torch.float64)
    good = ~torch.isnan(val)
    d = (j - i)
    bal_sum = torch.zeros(nb, dtype=torch.float64, device=dev).scatter_add_(0, d[good], val[good])
    okf = ok.to(torch.float64)
    nfft = 1 << int(np.ceil(np.log2(max(2 * nb, 2))))
    f = torch.fft.rfft(okf, nfft)
    n_valid = torch.round(torch.fft.irfft(f * torch.conj(f), nfft)[:nb])
    expected = bal_sum / n_valid
    expected[:2] = float("nan")
    raw_sum = torch.zeros(nb, dtype=torch.float64, device=dev).scatter_add_(0, d, c.to(torch.float64))
    expected_raw = raw_sum / torch.arange(nb, 0, -1, device=dev, dtype=torch.float64)
    expected_raw[:2] = float("nan")
    # the upper triangle as a cooler stores it (row-major, columns sorted): input of pup_region_create_upper
    up_indptr = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
    up_indptr[1:] = torch.cumsum(torch.bincount(i, minlength=nb), 0)
    upper = {"upper_indptr": up_indptr.to(torch.int32), "upper_col": j.to(torch.int32), "upper_count": c.clone()}
    # symmetric fill + sort into CSR
    off = i != j
    row = torch.cat([i, j[off]])
    col = torch.cat([j, i[off]])
    cnt = torch.cat([c, c[off]])
    del i, j, c, val, good, d
    key = row * nb + col
    del row, col
    key, order = torch.sort(key)
    cnt = cnt[order]
    del order
    row = torch.div(key, nb, rounding_mode="floor")
    col = (key - row * nb).to(torch.int32)
    del key
    indptr = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(torch.bincount(row, minlength=nb), 0)
    if int(indptr[-1]) >= 2**31:
        raise ValueError("region has more than 2^31 stored pixels")
    return {"nb": nb, "indptr": indptr.to(torch.int32), "col": col.contiguous(), "count": cnt.contiguous(),
            "weight": w, "expected": expected, "expected_raw": expected_raw, **upper}


def chrom_bins(chromsizes=None, binsize=10_000):
    chromsizes = HG38 if chromsizes is None else chromsizes
    return {c: -(-int(n) // binsize) for c, n in chromsizes.items()}


def synthetic_sites(n_pairs_target, chromsizes=None, binsize=10_000, flank=410_000, seed=1237):
    """CTCF-like stranded sites whose all-vs-all cis pairs passing ``mindist='auto'`` number ~``n_pairs_target``."""
    chromsizes = HG38 if chromsizes is None else chromsizes
    rng = np.random.default_rng(seed)
    total = float(sum(chromsizes.values()))
    mindist = 2 * flank + 2 * binsize

    def build(n_total):
        rows = []
        r = np.random.default_rng(seed)
        for c, L in chromsizes.items():
            n = max(2, int(round(n_total * L / total)))
            lo, hi = flank + binsize, L - flank - 2 * binsize
            pos = np.sort(r.integers(lo, hi, n))
            strand = r.choice(["+", "-"], n)
            rows.append(pd.DataFrame({"chrom": c, "start": pos, "end": pos + 200, "name": "site", "score": 0, "strand": strand}))
        return pd.concat(rows, ignore_index=True)

    def count_pairs(df):
        tot = 0
        for _, sub in df.groupby("chrom", sort=False):
            ctr = (sub["start"].values + sub["end"].values) / 2
            # pairs with |c2 - c1| >= mindist: total pairs minus close pairs
            hi_idx = np.searchsorted(ctr, ctr + mindist, side="left")
            tot += int((len(ctr) - hi_idx).sum())
        return tot

    lo_n, hi_n = 100, 200_000
    for _ in range(40):
        mid = (lo_n + hi_n) // 2
        if count_pairs(build(mid)) < n_pairs_target:
            lo_n = mid + 1
        else:
            hi_n = mid
    df = build(lo_n)
    del rng
    return df, count_pairs(df)


def synthetic_loops(n_loops, chromsizes=None, binsize=10_000, flank=410_000, seed=1236, dmin=1_000_000,
                    dmax=10_000_000):
    """bedpe loops: anchors uniform over the genome, separation log-uniform in [dmin, dmax]."""
    chromsizes = HG38 if chromsizes is None else chromsizes
    rng = np.random.default_rng(seed)
    names = list(chromsizes)
    sizes = np.array([chromsizes[c] for c in names], dtype=np.float64)
    ch = rng.choice(len(names), n_loops, p=sizes / sizes.sum())
    dist = np.exp(rng.uniform(np.log(dmin), np.log(dmax), n_loops)).astype(np.int64)
    L = sizes[ch].astype(np.int64)
    margin = flank + 2 * binsize
    span = np.maximum(L - dist - 2 * margin, 1)
    a = margin + (rng.random(n_loops) * span).astype(np.int64)
    b = a + dist
    keep = b + margin < L
    df = pd.DataFrame({"chrom1": np.array(names)[ch], "start1": a, "end1": a + binsize,
                       "chrom2": np.array(names)[ch], "start2": b, "end2": b + binsize})
    return df[keep].reset_index(drop=True)


def synthetic_cooler(chromsizes, binsize=10_000, depth=50.0, seed=0, device="cuda", nan_frac=0.03):
    """A :class:`~coolpuppy_b200.coolio.MemCooler` holding a small synthetic genome (cis pixels only) with a
    ``weight`` column, plus a matching expected table (``balanced.avg`` / ``count.avg``) for the whole-chromosome
    view.  Used by the mid-size parity tests of BASELINE configs[2] and configs[4]."""
    from .coolio import MemCooler

    bins = chrom_bins(chromsizes, binsize)
    b1, b2, cnt, ws, exp_rows = [], [], [], [], []
    off = 0
    for ci, (c, nb) in enumerate(bins.items()):
        t = synthetic_region(nb, depth=depth, seed=seed + ci, device=device, nan_frac=nan_frac)
        ip = t["upper_indptr"].cpu().numpy().astype(np.int64)
        rows = np.repeat(np.arange(nb, dtype=np.int64), np.diff(ip))
        b1.append(rows + off)
        b2.append(t["upper_col"].cpu().numpy().astype(np.int64) + off)
        cnt.append(t["upper_count"].cpu().numpy().astype(np.int32))
        ws.append(t["weight"].cpu().numpy())
        exp_rows.append(pd.DataFrame({"region1": c, "region2": c, "dist": np.arange(nb), "n_valid": nb - np.arange(nb),
                                      "balanced.avg": t["expected"].cpu().numpy(),
                                      "count.avg": t["expected_raw"].cpu().numpy()}))
        off += nb
    clr = MemCooler(chromsizes, binsize, np.concatenate(b1), np.concatenate(b2), np.concatenate(cnt),
                    {"weight": np.concatenate(ws)}, filename="synthetic.cool")
    return clr, pd.concat(exp_rows, ignore_index=True)

#!/usr/bin/env python
"""Benchmark of the pile-up hot path on B200 (contract: see the task brief, section 4).

Workloads (synthetic 3 Gbp genome at 10 kb bins, hg38 chromosome lengths, counts ~ Poisson(depth / separation),
balanced with 3 % NaN bins):
  configs3 (default; BASELINE.json configs[3], the configuration the metric is quoted on; it fits one GPU):
           all-vs-all cis pairs of CTCF-like sites chosen so that ~1e6 pairs pass mindist="auto", pad = 41 bins
           (W = 83), nshifts = 10 random-shift controls (~1.1e7 windows per step), no expected.
  configs2 (BASELINE.json configs[2]): 1e5 bedpe loops, pad = 41, observed / expected.
  configs4 (BASELINE.json configs[4]): ~5e5 stranded cis pairs, by-strand x by-distance groups, pad = 101 (W = 203),
           observed / expected, one accumulator slot per (strand1, strand2, distance band).

A step = one pass of the hot path over every window of every chromosome:
  value   : region matrices and window arrays resident in HBM; per sharding unit pup_accumulate() = device sort of the
            windows + count kernels + main pile-up kernel; N > 1: the (chromosome, row anchor) sequence is cut into N
            contiguous pieces of equal exact algorithmic bytes (a chromosome on a cut is piled up as row bands on two
            ranks, matrix replicated), one NCCL all-reduce of the accumulators inside the timed region.
  e2e     : the same pass through the product's two-stream region pipeline (coolpuppy_b200.pipeline) with HOST (pinned)
            buffers: pup_region_create_upper() + pup_upload() + pup_accumulate() per chromosome, i.e. including the
            H2D upload and device-side indexing of every chromosome and the D2H read of the accumulators.
  e2e_api : wall time of the public call coolpuppy_b200.coolpup.pileup(clr, features, ...) on the same workload --
            feature table in, DataFrame out; window generation (on the GPU), uploads, pile-up, export, normalisation.
  cpu_baseline / --impl reference : the restated reference path (oracle restatement of _stream_snips + _add_snip:
            scipy-CSR slice per window, NaN masks, expected divide, nansum) on the host cores, on a bounded sample of
            the same windows drawn in proportion to every chromosome's share of the windows.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BINSIZE = 10_000
WORKLOADS = {
    "configs3": dict(flank=410_000, nshifts=10, pairs=1_000_000, expected=False, features="sites", sites_seed=1237,
                     kwargs=dict(features_format="bed", nshifts=10, seed=0)),
    "configs2": dict(flank=410_000, nshifts=0, loops=100_000, expected=True, features="loops",
                     kwargs=dict(features_format="bedpe", nshifts=0, ooe=True)),
    "configs4": dict(flank=1_010_000, nshifts=0, pairs=500_000, expected=True, features="sites", sites_seed=1238,
                     kwargs=dict(features_format="bed", nshifts=0, ooe=True, by_strand=True, by_distance=True)),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PUP_BENCH_WORKLOAD", "configs3"), choices=list(WORKLOADS))
    ap.add_argument("--depth", type=float, default=float(os.environ.get("PUP_BENCH_DEPTH", 500.0)))
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("PUP_BENCH_PAIRS", 0)), help="override the pair target")
    ap.add_argument("--chroms", default=os.environ.get("PUP_BENCH_CHROMS", "all"), help="'all' or comma list (debug)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--api-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=30000, help="windows in the cpu_baseline sample (~10 s on one core)")
    ap.add_argument("--ref-sample", type=int, default=57600, help="windows per step over all chromosomes, --impl reference")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-api", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload
def chromsizes(args):
    from coolpuppy_b200.synthetic import HG38

    if args.chroms == "all":
        return dict(HG38)
    return {c: HG38[c] for c in args.chroms.split(",")}


def make_features(args, sizes):
    """Feature table of the workload (bed sites or bedpe loops) and the number of ROI pairs it yields."""
    from coolpuppy_b200.synthetic import HG38, synthetic_loops, synthetic_sites

    wl = WORKLOADS[args.workload]
    scale = float(sum(sizes.values())) / float(sum(HG38.values()))
    if wl["features"] == "loops":
        loops = synthetic_loops(int(round(wl["loops"] * scale)), chromsizes=sizes, binsize=BINSIZE, flank=wl["flank"], seed=1236)
        return loops, len(loops)
    target = int(round((args.pairs or wl["pairs"]) * scale))
    sites, n_pairs = synthetic_sites(target, chromsizes=sizes, binsize=BINSIZE, flank=wl["flank"], seed=wl["sites_seed"])
    return sites, n_pairs


def workload_description(args, n_features, W):
    wl = WORKLOADS[args.workload]
    if args.workload == "configs3":
        return (f"configs[3]: synthetic 3 Gbp @10 kb (hg38 lengths, Poisson(depth/sep)), all-vs-all cis pairs of {n_features} "
                f"CTCF-like sites, pad=41 (W={W}), nshifts={wl['nshifts']}, balanced (3% NaN bins), no expected")
    if args.workload == "configs2":
        return (f"configs[2]: synthetic 3 Gbp @10 kb (hg38 lengths, Poisson(depth/sep)), {n_features} bedpe loops (1-10 Mb), "
                f"pad=41 (W={W}), balanced (3% NaN bins), observed/expected")
    return (f"configs[4]: synthetic 3 Gbp @10 kb (hg38 lengths, Poisson(depth/sep)), all-vs-all cis pairs of {n_features} "
            f"stranded sites, by-strand x by-distance groups, pad=101 (W={W}), balanced (3% NaN bins), observed/expected")


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
_CPU = {}


def _cpu_worker_prepare(name):
    """Build the region's balanced scipy CSR once (untimed: cooler fetch + balancing in the reference)."""
    from scipy import sparse

    d = _CPU["regions"][name]
    nb = d["nb"]
    up = sparse.csr_matrix((d["upper_count"].astype(np.float64), d["upper_col"], d["upper_indptr"]), shape=(nb, nb))
    mat = (up + sparse.triu(up, 1).T).tocoo()  # cooler's fetch mirrors the stored upper triangle
    w = d["weight"]
    mat.data = w[mat.row] * w[mat.col] * mat.data
    d["mat"] = mat.tocsr()
    d["isnan"] = np.isnan(w)
    return name


def _cpu_worker_step(name):
    """Restated _stream_snips + _add_snip over the region's sample of windows (coolpup.py:1104-1157, puputils 12-38)."""
    d = _CPU["regions"][name]
    if "mat" not in d:
        _cpu_worker_prepare(name)
    mat, isnan, nb, W = d["mat"], d["isnan"], d["nb"], _CPU["W"]
    E = d.get("expected")
    acc = {}
    n = 0
    ii0 = np.arange(W)[:, None]
    jj0 = np.arange(W)[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        for s1, s2, k in zip(d["r0"], d["c0"], d["slot"]):
            if s1 < 0 or s1 + W > nb or s2 < 0 or s2 + W > nb:
                continue
            data = mat[s1 : s1 + W, s2 : s2 + W].toarray().astype(float)
            data[isnan[s1 : s1 + W], :] = np.nan
            data[:, isnan[s2 : s2 + W]] = np.nan
            dd = (s2 + jj0) - (s1 + ii0)
            data[dd < 2] = np.nan
            if E is not None:
                data = data / E[np.abs(dd)]
            if k not in acc:
                acc[k] = [data, np.isfinite(data).astype(int)]
            else:
                a = acc[k]
                a[0] = np.nansum([a[0], data], axis=0)
                a[1] += np.isfinite(data).astype(int)
            n += 1
    return n


def cpu_setup(regions_host, windows, total_sample, W, seed=99):
    """A sample of the workload's windows for the CPU arms: every region contributes in proportion to its share of
    the windows (uniformly drawn inside the region)."""
    rng = np.random.default_rng(seed)
    n_all = sum(len(windows[c]["r0"]) for c in regions_host)
    regs = {}
    for name, d in regions_host.items():
        w = windows[name]
        n = len(w["r0"])
        k = min(n, int(round(total_sample * n / max(1, n_all))))
        idx = np.sort(rng.choice(n, k, replace=False)) if k else np.zeros(0, dtype=np.int64)
        regs[name] = dict(d, r0=w["r0"][idx], c0=w["c0"][idx], slot=w["slot"][idx])
    _CPU["regions"] = regs
    _CPU["W"] = W


def run_cpu_pool(names, nproc, steps, warmup):
    """One process per region task like Pool.starmap over regions (coolpup.py:1502-1508); returns windows/s."""
    import multiprocessing as mp

    from coolpuppy_b200.multigpu import lpt_assign

    ctx = mp.get_context("fork")
    names = [n for n in names if len(_CPU["regions"][n]["r0"])]
    nproc = max(1, min(nproc, len(names)))
    owner = lpt_assign([len(_CPU["regions"][n]["r0"]) * max(1.0, _CPU["regions"][n]["nb"] / 1e4) for n in names], nproc)
    groups = [[n for n, o in zip(names, owner) if o == r] for r in range(nproc)]

    def worker(conn, mine):
        for n in mine:
            _cpu_worker_prepare(n)
        conn.send("ready")
        while True:
            msg = conn.recv()
            if msg == "stop":
                return
            conn.send(sum(_cpu_worker_step(n) for n in mine))

    procs = []
    for mine in groups:
        a, b = ctx.Pipe()
        p = ctx.Process(target=worker, args=(b, mine), daemon=True)
        p.start()
        procs.append((p, a))
    for _, c in procs:
        assert c.recv() == "ready"
    total = 0
    for _ in range(warmup):
        for _, c in procs:
            c.send("step")
        for _, c in procs:
            c.recv()
    t0 = time.perf_counter()
    for _ in range(steps):
        for _, c in procs:
            c.send("step")
        total += sum(c.recv() for _, c in procs)
    dt = time.perf_counter() - t0
    for p, c in procs:
        c.send("stop")
        p.join(timeout=10)
    return total / dt, dt / steps, total // max(steps, 1), nproc


# ------------------------------------------------------------------------------------------------ genome + windows
def generate_genome(args, sizes, dev, keep_device, keep_host, pin):
    """Per chromosome: the cooler-style upper triangle (+ weights, expected) as device tensors and / or host arrays."""
    import torch

    from coolpuppy_b200.synthetic import synthetic_region

    dev_data, host_data = {}, {}
    for ci, (c, L) in enumerate(sizes.items()):
        nb = -(-L // BINSIZE)
        t = synthetic_region(nb, depth=args.depth, seed=1234 + ci, device=dev, nan_frac=0.03)
        keys = ("upper_indptr", "upper_col", "upper_count", "weight", "expected")
        if c in keep_device:
            dev_data[c] = {k: t[k] for k in keys}
            dev_data[c]["nb"] = nb
        if c in keep_host:
            h = {}
            for k in keys:
                x = t[k].cpu()
                h[k] = x.pin_memory() if pin else x
            h["nb"] = nb
            host_data[c] = h
        del t
    if torch.device(dev).type == "cuda":
        torch.cuda.synchronize(dev)
        torch.cuda.empty_cache()
    return dev_data, host_data


def ensure_fast_pinned(host_data, dev, threshold_gbs=45.0, rounds=3):
    """Pinned host buffers whose plain H2D copy runs well below the link rate (seen intermittently on the VM hosts of
    this pool: a whole process at ~33 GB/s instead of ~55) are re-allocated: the old block stays allocated until the
    end so that the allocator hands out different memory.  Returns (arrays re-pinned, arrays still slow, slowest
    rate, the replaced blocks -- keep them referenced until the run ends).  Input staging only: nothing here is part of
    a timed region."""
    import torch

    def rate(t):
        d = torch.empty_like(t, device=dev)
        d.copy_(t, non_blocking=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        d.copy_(t, non_blocking=True)
        b.record()
        torch.cuda.synchronize(dev)
        return t.numel() * t.element_size() / 1e9 / (a.elapsed_time(b) / 1e3)

    keep, repinned, slow, worst = [], 0, 0, float("inf")
    for c, h in host_data.items():
        for k, t in list(h.items()):
            if not hasattr(t, "is_pinned") or not t.is_pinned() or t.numel() * t.element_size() < (32 << 20):
                continue
            r = rate(t)
            for _ in range(rounds):
                if r >= threshold_gbs:
                    break
                keep.append(t)
                t = torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
                h[k] = t
                repinned += 1
                r = rate(t)
            slow += int(r < threshold_gbs)
            worst = min(worst, r)
    return repinned, slow, (None if worst == float("inf") else worst), keep


def expected_table(host_or_dev):
    import pandas as pd

    rows = []
    for c, d in host_or_dev.items():
        e = d["expected"]
        e = e.cpu().numpy() if hasattr(e, "cpu") else np.asarray(e)
        nb = d["nb"]
        rows.append(pd.DataFrame({"region1": c, "region2": c, "dist": np.arange(nb), "n_valid": nb - np.arange(nb),
                                  "balanced.avg": e}))
    return pd.concat(rows, ignore_index=True)


def host_cooler(sizes, host_data, pin):
    """The workload's genome as a cooler object (coolio.ChromCooler: per-chromosome upper-triangle CSR)."""
    from coolpuppy_b200.coolio import ChromCooler

    regions = {c: tuple(np.asarray(h[k]) if not hasattr(h[k], "numpy") else h[k].numpy()
                        for k in ("upper_indptr", "upper_col", "upper_count")) for c, h in host_data.items()}
    nb = {c: -(-L // BINSIZE) for c, L in sizes.items()}
    weight = np.concatenate([(host_data[c]["weight"].numpy() if c in host_data else np.full(nb[c], np.nan)) for c in sizes])
    clr = ChromCooler(sizes, BINSIZE, regions, {"weight": weight}, filename="synthetic_3Gbp_10kb.cool", pin=pin)
    return clr


def build_windows(args, sizes, features, expected_df, nnz_by_chrom=None):
    """Host side of the path through the product's own builder (PileUpper._prepare): per-chromosome window arrays
    (reference emission order, seeded control shifts) with their dense accumulator slots."""
    from functools import partial

    from coolpuppy_b200 import coolpup as cp
    from coolpuppy_b200.coolio import ChromCooler

    wl = WORKLOADS[args.workload]
    kw = wl["kwargs"]
    nbins = sum(-(-L // BINSIZE) for L in sizes.values())
    clr = ChromCooler(sizes, BINSIZE, {}, {"weight": np.ones(nbins)}, filename="synthetic_3Gbp_10kb.cool")
    if nnz_by_chrom:  # stored pixels per chromosome: the density scale of the product's sharding cost model
        edges = np.concatenate([[0.0], np.cumsum([float(nnz_by_chrom[c]) for c in sizes])])
        clr._bin1_offset = np.interp(np.arange(nbins + 1), clr._chrom_offset, edges).astype(np.int64)
    np.random.seed(0)
    cc = cp.CoordCreator(features, BINSIZE, features_format=kw["features_format"], flank=wl["flank"], nshifts=wl["nshifts"],
                         mindist="auto", seed=0 if wl["nshifts"] else None)
    pu = cp.PileUpper(clr, cc, clr_weight_name="weight", expected=expected_df if wl["expected"] else False, ooe=True,
                      control=wl["nshifts"] > 0)
    groupby, modify = [], None
    if kw.get("by_strand"):
        groupby += ["strand1", "strand2"]
    if kw.get("by_distance"):
        modify = partial(cp.bin_distance_intervals, band_edges=pu._distance_edges("default"))
        groupby += ["distance_band"]
    plan = pu._plan(groupby, False, modify, None)
    plan["band_edges"] = pu._band_edges(plan)
    job = pu._prepare(plan, None, None)
    out = {c: dict(nb=-(-L // BINSIZE), r0=np.zeros(0, np.int32), c0=np.zeros(0, np.int32), slot=np.zeros(0, np.int32), n_roi=0,
                   anchor=np.zeros(0, np.int64)) for c, L in sizes.items()}
    for b in job["built"]:
        out[b["name"]].update(r0=b["w_r0"].astype(np.int32), c0=b["w_c0"].astype(np.int32), slot=b["slot"].astype(np.int32),
                              n_roi=int(np.count_nonzero(b["valid"] & (b["rw"].kind == 0))),
                              anchor=np.repeat(np.asarray(b["rw"].idx1), b["targets"]))
    return out, job["n_slots"], job["flags"], pu


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return 0 if rank != 0 else run_reference(args)

    import torch

    from coolpuppy_b200 import _native
    from coolpuppy_b200.multigpu import contiguous_partition, lpt_assign

    _native.require_device()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group(backend="nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    W = 2 * (wl["flank"] // BINSIZE) + 1
    sizes = chromsizes(args)
    names = list(sizes)
    features, n_pairs = make_features(args, sizes)
    want_host = (not args.no_e2e) or (not args.no_api) or (rank == 0 and not args.no_cpu and world == 1)
    # every rank generates every chromosome (seeded, identical): exact per-chromosome bytes need the matrices
    dev_data, host_data = generate_genome(args, sizes, dev, keep_device=set(names), keep_host=set(names) if want_host else set(),
                                          pin=world == 1)
    expected_df = expected_table(dev_data) if wl["expected"] else None
    pin_report, _pin_keep = None, None
    if world == 1 and want_host and os.environ.get("PUP_BENCH_REPIN", "1") != "0":
        n_re, n_slow, worst, _pin_keep = ensure_fast_pinned(host_data, dev)
        pin_report = {"arrays_repinned": n_re, "arrays_still_slow": n_slow, "slowest_pinned_h2d_gbs": worst}
    t_host0 = time.perf_counter()
    nnz_by_chrom = {c: int(dev_data[c]["upper_col"].shape[0]) for c in names}
    windows, n_slots, flags, pu_plan = build_windows(args, sizes, features, expected_df, nnz_by_chrom)
    host_window_s = time.perf_counter() - t_host0
    n_windows_total = sum(len(w["r0"]) for w in windows.values())
    n_roi_total = sum(w["n_roi"] for w in windows.values())
    region_flags = flags & _native.PUP_F_OOE
    acc_flags = flags & (_native.PUP_F_EXPCTRL | _native.PUP_F_COVERAGE)

    # ---- resident regions + exact algorithmic bytes of every chromosome's windows (identical on all ranks)
    stream = torch.cuda.current_stream(dev).cuda_stream
    regions, cost, nnz_win = {}, [], []
    for c in names:
        d = dev_data[c]
        regions[c] = _native.Region(local_rank, d["nb"], d["upper_indptr"], d["upper_col"], d["upper_count"], d["weight"],
                                    d["expected"] if wl["expected"] else None, None, ignore_diags=2, flags=region_flags,
                                    stream=stream, upper=True)
        w = windows[c]
        if len(w["r0"]):
            b, z = regions[c].algorithmic_bytes(torch.from_numpy(w["r0"]).to(dev), torch.from_numpy(w["c0"]).to(dev), W, flags,
                                                stream=stream)
        else:
            b, z = 0, 0
        cost.append(b)
        nnz_win.append(z)
    if os.environ.get("PUP_BENCH_VERBOSE") and rank == 0:
        sys.stderr.write(json.dumps({"algorithmic_bytes_by_chromosome": dict(zip(names, [int(c) for c in cost])),
                                     "windows_by_chromosome": {c: len(windows[c]["r0"]) for c in names}}) + "\n")
    # one contiguous, equal-cost piece of the (chromosome, row anchor) sequence per rank (the product's sharding,
    # multigpu.contiguous_partition) with the product's predicted pile-up time per row anchor (PileUpper._feature_costs)
    arrs = []
    for c, b in zip(names, cost):  # the product's own time model (sparse / dense-band kernel), not rescaled
        arrs.append(np.array(pu_plan._feature_costs(c) if len(windows[c]["r0"]) else np.zeros(0), dtype=np.float64))
    pieces, loads = contiguous_partition(arrs, world)
    my_units = [(names[i], lo, hi, lo == 0 and hi == len(arrs[i])) for i, lo, hi in pieces[rank]]
    predicted_imbalance = float(max(loads) / (sum(loads) / len(loads))) if sum(loads) > 0 else 1.0
    for c in names:  # matrices of chromosomes this rank has no piece of are not needed any more
        if c not in {u[0] for u in my_units}:
            regions.pop(c).close()
    split = [names[i] for i in sorted({i for p_ in pieces for i, lo, hi in p_ if not (lo == 0 and hi == len(arrs[i]))})]
    n_units = sum(len(p_) for p_ in pieces)
    dwin = {}
    alg_bytes = 0
    for c, lo, hi, whole in my_units:
        w = windows[c]
        sel = np.arange(len(w["r0"])) if whole else np.nonzero((w["anchor"] >= lo) & (w["anchor"] < hi))[0]
        dwin[(c, lo)] = tuple(torch.from_numpy(np.ascontiguousarray(w[k][sel])).to(dev) for k in ("r0", "c0", "slot"))
        if whole:
            alg_bytes += cost[names.index(c)]
        elif len(sel):
            alg_bytes += regions[c].algorithmic_bytes(dwin[(c, lo)][0], dwin[(c, lo)][1], W, flags, stream=stream)[0]
    if not want_host:
        host_data = {}
    del dev_data
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()

    config = {
        "workload": workload_description(args, len(features), W),
        "depth": args.depth, "roi_windows": int(n_pairs), "windows_per_step": int(n_windows_total),
        "chromosomes": len(names), "binsize": BINSIZE, "flank": wl["flank"], "accumulator_slots": int(n_slots),
        "l2": "inputs (region matrices, GBs) exceed the 126 MB L2; no flush between iterations",
        "parallelism": (f"the (chromosome, row anchor) sequence is cut into {world} contiguous pieces of equal predicted pile-up time, one "
                        "per GPU (a chromosome that straddles a cut is piled up as row bands on two GPUs, its matrix "
                        "replicated); one all-reduce of the accumulators; e2e: whole chromosomes per GPU by LPT"),
    }

    stride = _native.acc_stride(W)
    acc = torch.zeros(n_slots * stride, dtype=torch.float64, device=dev)

    def step():
        acc.zero_()
        launches = 1
        for c, lo, _, _ in my_units:
            r0, c0, sl = dwin[(c, lo)]
            if r0.shape[0] == 0:
                continue
            regions[c].accumulate(r0, c0, sl, W, n_slots, acc_flags, acc, stream=stream)
            launches += _native.lib().pup_last_launches()
        if dist is not None:
            dist.all_reduce(acc)
        return launches

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    _native.timing_enable(True)
    _native.timing_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        launches += step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    phases = _native.timing_read(reset=True)
    _native.timing_enable(False)
    clocks = sampler.stop()
    # all-reduce time alone (payload = the accumulators), N > 1
    allreduce_ms = None
    if dist is not None:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a0.record()
        for _ in range(5):
            dist.all_reduce(acc)
        a1.record()
        barrier()
        allreduce_ms = a0.elapsed_time(a1) / 5
        step()  # restore the accumulators of one clean step
        barrier()

    n_all = _native.acc_counts(acc, W, n_slots)
    n_valid_total = int(n_all.sum())  # after the all-reduce: whole job
    sums = acc.view(n_slots, stride)[:, : W * W]
    checksum = {"n": n_valid_total, "sum": float(f"{float(sums[torch.isfinite(sums)].sum().item()):.10e}"),
                "slots_used": int((n_all > 0).sum())}

    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    stats = torch.tensor([alg_bytes, phases["main"][0] + phases["dense_band"][0], launches, phases["main"][1]], dtype=torch.float64,
                         device=dev)  # the pile-up phase = sparse kernel + dense-band kernel of every call
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        allstats = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allstats, stats)
    else:
        allstats = [stats]
    ms = float(t_ms.item())
    ms_per_step = ms / args.steps
    value = n_valid_total / (ms_per_step / 1e3)
    roi_valid = n_roi_total  # ROI windows that lie inside their region (controls not counted)

    # ---- e2e through the C ABI with host buffers: the product's region pipeline
    e2e = None
    cost_e2e = [c_ + 40.0 * (host_data[n]["upper_col"].numel() if n in host_data else 0) for n, c_ in zip(names, cost)]
    owner_e2e = lpt_assign(cost_e2e, world)
    mine = [c for c, o in zip(names, owner_e2e) if o == rank]
    if not args.no_e2e:
        from coolpuppy_b200.pipeline import RegionPipeline

        hacc = torch.zeros(n_slots * stride, dtype=torch.float64).pin_memory()
        if world > 1:  # this rank's chromosomes into pinned memory (N = 1: the whole genome already is)
            for c in mine:
                host_data[c] = {k: (v.pin_memory() if hasattr(v, "pin_memory") else v) for k, v in host_data[c].items()}
        hwin = {c: tuple(torch.from_numpy(windows[c][k]).pin_memory() for k in ("r0", "c0", "slot")) for c in mine}
        hk = ("upper_indptr", "upper_col", "upper_count", "weight") + (("expected",) if wl["expected"] else ())
        h2d = sum(sum(host_data[c][k].numel() * host_data[c][k].element_size() for k in hk) for c in mine)
        h2d += sum(sum(t.numel() * t.element_size() for t in hwin[c]) for c in mine)
        d2h = acc.numel() * 8
        # upload order: one small chromosome first (its upload is the only one nothing overlaps), then big to small
        e2e_order = sorted(mine, key=lambda c: -windows[c]["nb"])
        if len(e2e_order) > 2:
            e2e_order = [e2e_order[-1]] + e2e_order[:-1]

        def e2e_step():
            acc.zero_()
            pipe = RegionPipeline(local_rank, W, n_slots, acc_flags)
            for c in e2e_order:
                h = host_data[c]
                kw = dict(nb=h["nb"], indptr=h["upper_indptr"], col=h["upper_col"], count=h["upper_count"], weight=h["weight"],
                          expected=h["expected"] if wl["expected"] else None, coverage=None, ignore_diags=2,
                          flags=region_flags, upper=True)
                pipe.submit(kw, tuple(t.numpy() for t in hwin[c]), acc)
            pipe.finish()
            if dist is not None:
                dist.all_reduce(acc)
            hacc.copy_(acc, non_blocking=True)
            torch.cuda.synchronize(dev)

        e2e_step()
        barrier()
        # plain pinned-host -> device copy rate of this process's largest pixel array (the ceiling of the e2e step)
        probe_src = max((host_data[c]["upper_col"] for c in mine), key=lambda t: t.numel(), default=None)
        h2d_probe = None
        if probe_src is not None and probe_src.numel() > 0:
            probe_dst = torch.empty_like(probe_src, device=dev)
            probe_dst.copy_(probe_src, non_blocking=True)
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            probe_dst.copy_(probe_src, non_blocking=True)
            p1.record()
            torch.cuda.synchronize(dev)
            h2d_probe = probe_src.numel() * probe_src.element_size() / 1e9 / (p0.elapsed_time(p1) / 1e3)
            del probe_dst
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        step_wall = []
        for _ in range(args.e2e_steps):
            t_w = time.perf_counter()
            e2e_step()
            step_wall.append((time.perf_counter() - t_w) * 1e3)
        a1.record()
        barrier()
        if os.environ.get("PUP_BENCH_VERBOSE") and rank == 0:
            sys.stderr.write(json.dumps({"e2e_step_wall_ms": [round(x, 1) for x in step_wall]}) + "\n")
        ems = torch.tensor([a0.elapsed_time(a1) / args.e2e_steps], dtype=torch.float64, device=dev)
        bts = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            dist.all_reduce(bts)
        e2e = {"value": n_valid_total / (float(ems.item()) / 1e3), "unit": "pile-ups/s", "ms_per_step": float(ems.item()),
               "h2d_bytes_per_step": int(bts[0].item()), "d2h_bytes_per_step": int(bts[1].item()),
               "steps": args.e2e_steps, "h2d_probe_gbs": h2d_probe, "pinned_input_check": pin_report,
               "step_wall_ms_min_max": [min(step_wall), max(step_wall)],
               "what": "coolpuppy_b200.pipeline.RegionPipeline (the product's own region loop) with pinned HOST buffers: per "
                       "chromosome pup_region_create_upper(cooler-style upper-triangle pixels, weights[, expected]) + pup_upload"
                       "(window arrays) on a prepare stream while the previous chromosome's pup_accumulate runs on a compute "
                       "stream; D2H of the accumulators at the end"}

    # ---- e2e through the public API: pileup(clr, features, ...) -> DataFrame
    e2e_api = None
    if not args.no_api:
        from coolpuppy_b200 import coolpup as cp
        from coolpuppy_b200.multigpu import RegionSharder

        # N = 1: views of the pinned host arrays; N > 1: every rank pins the chromosomes it reads on first use (the
        # untimed first call), not the whole genome
        clr = host_cooler(sizes, host_data, pin=False if world == 1 else "lazy")
        sharder = RegionSharder() if dist is not None else None
        kw = dict(wl["kwargs"], flank=wl["flank"], clr_weight_name="weight", device=local_rank, dist=sharder)
        if wl["expected"]:
            kw["expected_df"] = expected_df
        import logging
        import warnings

        logging.getLogger("coolpuppy").setLevel(logging.WARNING)
        times, stats_api, n_api = [], None, None
        for i in range(args.api_steps + 1):
            barrier()
            t0 = time.perf_counter()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pups = cp.pileup(clr, features, **kw)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            if i > 0:  # the first call warms the allocator / module state
                times.append(dt)
            is_all = [isinstance(g, str) and g == "all" for g in pups["group"]]
            n_api = int(pups.loc[is_all, "n"].iloc[0]) if any(is_all) else None
            stats_api = dict(cp._LAST_STATS)
        t_api = torch.tensor([float(np.median(times))], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_api, op=dist.ReduceOp.MAX)
        sec = float(t_api.item())
        api_windows = int(stats_api.get("windows", 0))
        e2e_api = {"value": api_windows / sec, "unit": "pile-ups/s", "seconds_per_call": sec, "calls": len(times),
                   "windows_per_call": api_windows, "rows": int(len(pups)), "n_all": n_api,
                   "host_prepare_s": stats_api.get("host_prepare_s"), "gpu_phase_s": stats_api.get("gpu_phase_s"),
                   "gpu_share": (stats_api.get("gpu_phase_s") or 0.0) / sec if sec > 0 else None,
                   "device_windows": stats_api.get("device_windows"), "sharding_imbalance_predicted": stats_api.get("imbalance"),
                   "what": "wall time of coolpuppy_b200.coolpup.pileup(clr, features, **kwargs) -> DataFrame: feature table in, "
                           "window generation" + (" + MT19937 control shifts" if wl["nshifts"] else "") + " on the GPU when the "
                           "workload allows, region pipeline from host memory, all-reduce, export, final normalisation"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_pileup_main): algorithmic bytes / its CUDA-event time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s"
    per_rank = [(float(s[0]), float(s[1]) / args.steps, float(s[3]) / args.steps) for s in allstats]
    tot_bytes = sum(b for b, _, _ in per_rank)
    slow_b, slow_ms, slow_launches = max(per_rank, key=lambda x: x[1])
    achieved = (slow_b / 1e9) / (slow_ms / 1e3) if slow_ms > 0 else 0.0
    kernel_ms = [m_ for _, m_, _ in per_rank]
    bytes_rank = [b for b, _, _ in per_rank]
    # DRAM traffic of the dominant kernel: from the committed `ncu --set full` capture of this round (it cannot be
    # measured outside a profiler); reported per launch, scaled by algorithmic bytes from the captured launch
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", f"r2_traffic_{args.workload}.json")))
        traffic = int(tr["dram_bytes"] / tr["algorithmic_bytes"] * (slow_b / max(1.0, slow_launches)))
        traffic_src = tr.get("source")
    except Exception:
        pass
    roofline = {
        "kernel": "k_pileup_main + k_pileup_dense (the pile-up phase: sparse strips + dense diagonal band)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "traffic_source": traffic_src,
        "algorithmic_bytes_per_step": int(tot_bytes), "algorithmic_bytes_per_launch": int(slow_b / max(1.0, slow_launches)),
        "kernel_ms_per_step": slow_ms, "launches_per_step": int(slow_launches),
        "avg_launch_ms": slow_ms / max(1.0, slow_launches),
        "stored_pixels_in_windows_per_step": int(sum(nnz_win)),
        "phase_ms_per_step": {k: v[0] / args.steps for k, v in phases.items()},
        "kernel_ms_per_step_by_rank": kernel_ms,
    }
    sharding = {"units": n_units, "split_chromosomes": split, "predicted_max_over_mean": predicted_imbalance,
                "bytes_max_over_mean": float(max(bytes_rank) / (sum(bytes_rank) / len(bytes_rank))) if sum(bytes_rank) else 1.0,
                "kernel_ms_max_over_mean": float(max(kernel_ms) / (sum(kernel_ms) / len(kernel_ms))) if sum(kernel_ms) else 1.0,
                "allreduce_payload_bytes": int(acc.numel() * 8), "allreduce_ms": allreduce_ms}

    # ---- CPU baseline: restated reference path on a bounded sample, 1 core
    cpu = None
    if not args.no_cpu and world == 1:
        rh = {c: dict(nb=host_data[c]["nb"], **{k: host_data[c][k].numpy() for k in ("upper_indptr", "upper_col", "upper_count", "weight")},
                      **({"expected": host_data[c]["expected"].numpy()} if wl["expected"] else {})) for c in names}
        cpu_setup(rh, windows, args.cpu_sample, W)
        used = [c for c in names if len(_CPU["regions"][c]["r0"])]
        for c in used:
            _cpu_worker_prepare(c)
        t0 = time.perf_counter()
        nwin = sum(_cpu_worker_step(c) for c in used)
        dt = time.perf_counter() - t0
        cpu = {"value": nwin / dt, "unit": "pile-ups/s", "cores": 1, "kind": "port",
               "sample": f"{nwin} windows drawn from all {len(used)} chromosomes in proportion to their share of the "
                         f"workload's windows ({dt:.1f} s; matrix fetch+balancing untimed), oracle restatement of "
                         "_stream_snips+_add_snip",
               "host_cpus": os.cpu_count()}

    line = {
        "metric": "pile-ups/sec (1e6 ROIs, 10kb bins, pad=41)", "value": value, "unit": "pile-ups/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "clocks": clocks, "e2e": e2e, "e2e_api": e2e_api, "gpu_launches": int(sum(float(s[2]) for s in allstats)),
        "roofline": roofline, "cpu_baseline": cpu, "sharding": sharding, "checksum": checksum,
        "windows_accumulated_per_step": n_valid_total,
        "value_roi_only": roi_valid / (ms_per_step / 1e3),
        "host_window_generation_s": {"host_builder_for_resident_arrays": round(host_window_s, 3),
                                     "api_host_prepare": None if e2e_api is None else e2e_api["host_prepare_s"]},
    }
    if args.workload != "configs3":
        line["metric"] = f"pile-ups/sec ({args.workload} of BASELINE.json)"
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_reference(args):
    """--impl reference: the restated reference CPU path on all usable host cores, bounded sample per step.

    Nothing of the product runs here: libpileup_b200.so is not loaded, windows are enumerated with numpy, the timed
    loop is numpy / scipy in forked worker processes.  The synthetic genome itself is generated with torch (on the GPU
    when one is visible -- untimed input preparation, labelled in the output -- else on the CPU)."""
    import torch

    wl = WORKLOADS[args.workload]
    W = 2 * (wl["flank"] // BINSIZE) + 1
    sizes = chromsizes(args)
    names = list(sizes)
    features, n_pairs = make_features(args, sizes)
    gen_dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    _, host_data = generate_genome(args, sizes, gen_dev, keep_device=set(), keep_host=set(names), pin=False) \
        if gen_dev.type == "cuda" else _generate_genome_cpu(args, sizes)
    windows = reference_windows(features, sizes, wl, W)
    n_windows_total = sum(len(w["r0"]) for w in windows.values())
    rh = {c: dict(nb=host_data[c]["nb"], **{k: np.asarray(host_data[c][k]) for k in ("upper_indptr", "upper_col", "upper_count", "weight")},
                  **({"expected": np.asarray(host_data[c]["expected"])} if wl["expected"] else {})) for c in names}
    cpu_setup(rh, windows, args.ref_sample, W)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    value, sec_per_step, per_step, nproc = run_cpu_pool(names, cores, args.steps, args.warmup)
    sample = (f"{per_step} windows per step drawn from all chromosomes in proportion to their share of the {n_windows_total} "
              f"windows, one process per chromosome task on {nproc} of {cores} cores like Pool.starmap over regions "
              "(coolpup.py:1502-1508); matrix fetch+balancing untimed")
    config = {
        "workload": workload_description(args, len(features), W),
        "depth": args.depth, "roi_windows": int(n_pairs), "windows_per_step": int(n_windows_total),
        "chromosomes": len(names), "binsize": BINSIZE, "flank": wl["flank"],
    }
    line = {
        "impl": "reference", "metric": "pile-ups/sec (1e6 ROIs, 10kb bins, pad=41)", "value": value, "unit": "pile-ups/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": "pile-ups/s", "cores": nproc, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pile-ups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "prep": f"synthetic genome generated with torch on {gen_dev.type} (untimed input preparation, not part of any arm); "
                "windows enumerated with numpy; libpileup_b200.so not loaded; timed loop = numpy/scipy on host cores",
    }
    if args.workload != "configs3":
        line["metric"] = f"pile-ups/sec ({args.workload} of BASELINE.json)"
    print(json.dumps(line))
    return 0


def _generate_genome_cpu(args, sizes):
    return generate_genome(args, sizes, "cpu", keep_device=set(), keep_host=set(sizes), pin=False)


def reference_windows(features, sizes, wl, W):
    """numpy enumeration of the workload's windows for the CPU arm (no native code): ROI pairs / loops per chromosome and
    their nshifts randomly shifted controls (coolpup.py:682-714, 387-453; same distribution as the product's seeded
    stream, not the same draws -- the CPU arm times a sample of them)."""
    res, flank, ns = BINSIZE, wl["flank"], wl["nshifts"]
    mindist = 2 * flank + 2 * res
    rng = np.random.default_rng(0)
    out = {}
    for c, L in sizes.items():
        nb = -(-L // res)
        if wl["features"] == "loops":
            f = features[features["chrom1"] == c]
            c1 = (f["start1"].values + f["end1"].values) / 2
            c2 = (f["start2"].values + f["end2"].values) / 2
            keep = np.abs(c2 - c1) >= mindist
            a = np.floor(c1[keep] / res).astype(np.int64) - flank // res
            b = np.floor(c2[keep] / res).astype(np.int64) - flank // res
        else:
            f = features[features["chrom"] == c].sort_values("start")
            ctr = (f["start"].values + f["end"].values) / 2
            st = np.floor(ctr / res).astype(np.int64) - flank // res
            k, l = np.triu_indices(len(ctr), 1)
            keep = np.abs(ctr[l] - ctr[k]) >= mindist
            a, b = st[k[keep]], st[l[keep]]
        slot = np.zeros(len(a), dtype=np.int32)
        if ns > 0 and len(a):
            sh = np.round(rng.integers(10**5, 10**6, len(a) * ns) * rng.choice([-1, 1], len(a) * ns) / res).astype(np.int64)
            a = np.concatenate([a, np.tile(a, ns) + sh])
            b = np.concatenate([b, np.tile(b, ns) + sh])
            slot = np.concatenate([slot, np.ones(len(sh), dtype=np.int32)])
        out[c] = dict(nb=nb, r0=a.astype(np.int32), c0=b.astype(np.int32), slot=slot)
    return out


if __name__ == "__main__":
    sys.exit(main())

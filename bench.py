#!/usr/bin/env python
"""Benchmark of the pile-up hot path on B200 (contract: see the task brief, section 4).

Workload (BASELINE.json configs[3], the configuration the metric is quoted on; it fits one GPU): synthetic 3 Gbp
genome at 10 kb bins (hg38 chromosome lengths, counts ~ Poisson(depth / separation)), all-vs-all cis pairs of
CTCF-like sites chosen so that ~1e6 pairs pass mindist="auto", pad = 41 bins (W = 83), nshifts = 10 random-shift
controls (~1.1e7 windows per step), balanced with 3 % NaN bins, no expected.

A step = one pass of the hot path over every window of every chromosome:
  value : region matrices and window arrays resident in HBM; per chromosome pup_accumulate() = device sort of the
          windows + vector kernel + main pile-up kernel; N > 1: chromosomes sharded over ranks (LPT), one NCCL
          all-reduce of the accumulators inside the timed region.
  e2e   : the same pass through pup_region_create_upper() + pup_upload() + pup_accumulate() with HOST (pinned)
          upper-triangle CSR / weight / window buffers, i.e. including the H2D upload + device-side indexing of every
          chromosome and the D2H read of the accumulators.
  cpu_baseline / --impl reference : the restated reference path (oracle/pileup_oracle.py: scipy-CSR slice per
          window, NaN masks, nansum) on the host cores, on a bounded uniform sample of the same windows.
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

CPU_SAMPLE_REGIONS = ("chr2", "chr9", "chr16", "chr21")
BINSIZE = 10_000
FLANK = 410_000
W = 2 * (FLANK // BINSIZE) + 1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--depth", type=float, default=float(os.environ.get("PUP_BENCH_DEPTH", 500.0)))
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("PUP_BENCH_PAIRS", 1_000_000)))
    ap.add_argument("--nshifts", type=int, default=10)
    ap.add_argument("--chroms", default=os.environ.get("PUP_BENCH_CHROMS", "all"), help="'all' or comma list (debug)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=30000, help="windows in the cpu_baseline sample (~10 s on one core)")
    ap.add_argument("--ref-sample", type=int, default=2400, help="windows per chromosome per step, --impl reference")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload
def chromsizes(args):
    from coolpuppy_b200.synthetic import HG38

    if args.chroms == "all":
        return dict(HG38)
    return {c: HG38[c] for c in args.chroms.split(",")}


def build_windows(args, sizes):
    """Host side of the path: features -> per-chromosome window arrays (reference order, seeded control shifts)."""
    from coolpuppy_b200.coolpup import CoordCreator
    from coolpuppy_b200.synthetic import synthetic_sites

    total = float(sum(sizes.values()))
    from coolpuppy_b200.synthetic import HG38

    target = int(round(args.pairs * (total / float(sum(HG38.values()))) ** 1))
    sites, n_pairs = synthetic_sites(target, chromsizes=sizes, binsize=BINSIZE, flank=FLANK, seed=1237)
    np.random.seed(0)
    cc = CoordCreator(sites, BINSIZE, features_format="bed", flank=FLANK, nshifts=args.nshifts, mindist="auto", seed=0)
    out = {}
    for c, L in sizes.items():
        rw = cc.region_windows((c, 0, L), control=args.nshifts > 0)
        nb = -(-L // BINSIZE)
        out[c] = dict(nb=nb, r0=rw.st1.astype(np.int32), c0=rw.st2.astype(np.int32), slot=rw.kind.astype(np.int32))
    return out, len(sites), n_pairs


def lpt(costs, n):
    from coolpuppy_b200.multigpu import lpt_assign

    return lpt_assign(costs, n)


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
    mat = sparse.csr_matrix((d["count"].astype(np.float64), d["col"], d["indptr"]), shape=(d["nb"], d["nb"]))
    w = d["weight"]
    coo = mat.tocoo()
    coo.data = w[coo.row] * w[coo.col] * coo.data
    d["mat"] = coo.tocsr()
    d["isnan"] = np.isnan(w)
    return name


def _cpu_worker_step(name):
    """Restated _stream_snips + _add_snip over the region's sample of windows (coolpup.py:1104-1157, puputils 12-38)."""
    d = _CPU["regions"][name]
    if "mat" not in d:
        _cpu_worker_prepare(name)
    mat, isnan, nb = d["mat"], d["isnan"], d["nb"]
    acc = {}
    n = 0
    ii0 = np.arange(W)[:, None]
    jj0 = np.arange(W)[None, :]
    for s1, s2, k in zip(d["r0"], d["c0"], d["slot"]):
        if s1 < 0 or s1 + W > nb or s2 < 0 or s2 + W > nb:
            continue
        data = mat[s1 : s1 + W, s2 : s2 + W].toarray().astype(float)
        data[isnan[s1 : s1 + W], :] = np.nan
        data[:, isnan[s2 : s2 + W]] = np.nan
        data[((s2 + jj0) - (s1 + ii0)) < 2] = np.nan
        if k not in acc:
            acc[k] = [data, np.isfinite(data).astype(int)]
        else:
            a = acc[k]
            a[0] = np.nansum([a[0], data], axis=0)
            a[1] += np.isfinite(data).astype(int)
        n += 1
    return n


def cpu_setup(regions_host, windows, per_region, seed=99):
    """Uniform sample of each region's windows for the CPU arms."""
    rng = np.random.default_rng(seed)
    regs = {}
    for name, d in regions_host.items():
        w = windows[name]
        n = len(w["r0"])
        k = min(per_region, n)
        idx = np.sort(rng.choice(n, k, replace=False)) if k else np.zeros(0, dtype=np.int64)
        regs[name] = dict(d, r0=w["r0"][idx], c0=w["c0"][idx], slot=w["slot"][idx])
    _CPU["regions"] = regs


def run_cpu_pool(names, nproc, steps, warmup):
    """One process per region task like Pool.starmap over regions (coolpup.py:1502-1508); returns windows/s."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    nproc = max(1, min(nproc, len(names)))
    owner = lpt([len(_CPU["regions"][n]["r0"]) * max(1.0, _CPU["regions"][n]["nb"] / 1e4) for n in names], nproc)
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


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0

    import torch

    from coolpuppy_b200 import _native
    from coolpuppy_b200.synthetic import synthetic_region

    _native.require_device()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1 and args.impl == "b200":
        import torch.distributed as dist

        dist.init_process_group(backend="nccl", device_id=dev)

    sizes = chromsizes(args)
    names = list(sizes)
    t_host0 = time.perf_counter()
    windows, n_sites, n_pairs = build_windows(args, sizes)
    host_window_s = time.perf_counter() - t_host0
    n_windows_total = sum(len(w["r0"]) for w in windows.values())
    config = {
        "workload": "configs[3]: synthetic 3 Gbp @10 kb (hg38 lengths, Poisson(depth/sep)), all-vs-all cis pairs of "
                    f"{n_sites} CTCF-like sites, pad=41 (W=83), nshifts={args.nshifts}, balanced (3% NaN bins), no expected",
        "depth": args.depth, "roi_windows": int(n_pairs), "windows_per_step": int(n_windows_total),
        "chromosomes": len(names), "binsize": BINSIZE, "flank": FLANK,
        "l2": "inputs (region matrices, GBs) exceed the 126 MB L2; no flush between iterations",
        "parallelism": f"chromosomes sharded over {world} GPU(s) by LPT, one all-reduce of the accumulators",
        "host_window_generation_s": round(host_window_s, 3),
    }

    # which chromosomes are mine (the sharding unit of the reference: one view region per worker, coolpup.py:1502-1508)
    cost = [len(windows[c]["r0"]) * (windows[c]["nb"] / 1e4) for c in names]
    owner = lpt(cost, world if args.impl == "b200" else 1)
    mine = [c for c, o in zip(names, owner) if o == (rank if args.impl == "b200" else 0)]
    # Resident pass, N > 1: a chromosome heavier than half a rank's share is split by WINDOWS into equal parts that go
    # to different ranks with the matrix replicated (windows are independent, the accumulators add up);
    # otherwise chr1 alone (13 % of the work) caps 8 ranks at 7.6x.  The e2e pass keeps whole chromosomes per rank:
    # there the matrix upload, not the pile-up, is the cost of a unit.
    from coolpuppy_b200.multigpu import part_bounds, split_heavy

    iunits, ucost, uowner = split_heavy(cost, world if args.impl == "b200" else 1)
    units = [(names[i], part, parts) for i, part, parts in iunits]
    my_units = [u for u, o in zip(units, uowner) if o == (rank if args.impl == "b200" else 0)]
    split = sorted({c for c, _, n in units if n > 1}, key=names.index)
    if split:
        config["parallelism"] = (f"chromosomes sharded over {world} GPUs by LPT; {','.join(split)} split by windows into "
                                 f"{sum(1 for c, _, n in units if n > 1)} parts (matrix replicated); one all-reduce of the "
                                 "accumulators; e2e: whole chromosomes per GPU")

    if args.impl == "reference":
        return run_reference(args, names, sizes, windows, config, dev)

    # ---- resident data
    stream = torch.cuda.current_stream(dev).cuda_stream
    regions, host, dwin = {}, {}, {}
    nnz_total = 0
    unit_chroms = {c for c, _, _ in my_units}
    for ci, c in enumerate(names):
        if c not in mine and c not in unit_chroms:
            continue
        t = synthetic_region(windows[c]["nb"], depth=args.depth, seed=1234 + ci, device=dev, nan_frac=0.03)
        if c in unit_chroms:
            nnz_total += int(t["col"].shape[0])
            regions[c] = _native.Region(local_rank, t["nb"], t["indptr"], t["col"], t["count"], t["weight"], None, None,
                                        ignore_diags=2, flags=0, stream=stream)
        if c in mine and (not args.no_e2e or (rank == 0 and not args.no_cpu)):
            host[c] = {k: t[k].cpu().pin_memory() for k in ("upper_indptr", "upper_col", "upper_count", "weight")}
            if rank == 0 and not args.no_cpu and c in CPU_SAMPLE_REGIONS:
                host[c].update({k: t[k].cpu() for k in ("indptr", "col", "count")})
        del t
    for c, part, parts in my_units:
        w = windows[c]
        lo, hi = part_bounds(len(w["r0"]), part, parts)
        dwin[(c, part)] = tuple(torch.from_numpy(np.ascontiguousarray(w[k][lo:hi])).to(dev) for k in ("r0", "c0", "slot"))
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()

    n_slots = 2
    stride = _native.acc_stride(W)
    acc = torch.zeros(n_slots * stride, dtype=torch.float64, device=dev)
    flags = 0

    # exact algorithmic bytes of my windows (untimed)
    alg_bytes = 0
    alg_nnz = 0
    for c, part, _ in my_units:
        b, z = regions[c].algorithmic_bytes(dwin[(c, part)][0], dwin[(c, part)][1], W, flags, stream=stream)
        alg_bytes += b
        alg_nnz += z

    def step():
        acc.zero_()
        launches = 1
        for c, part, _ in my_units:
            r0, c0, sl = dwin[(c, part)]
            regions[c].accumulate(r0, c0, sl, W, n_slots, flags, acc, stream=stream)
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

    out = _native.acc_export(acc, W, n_slots, device=local_rank, stream=stream)
    n_valid_total = int(out["n"].sum())  # after the all-reduce: whole job

    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    stats = torch.tensor([alg_bytes, alg_nnz, phases["main"][0], nnz_total, launches], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        allstats = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allstats, stats)
    else:
        allstats = [stats]
    ms = float(t_ms.item())
    ms_per_step = ms / args.steps
    value = n_valid_total / (ms_per_step / 1e3)

    # ---- e2e through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        hacc = torch.zeros(n_slots * stride, dtype=torch.float64).pin_memory()
        hwin = {c: tuple(torch.from_numpy(windows[c][k]).pin_memory() for k in ("r0", "c0", "slot")) for c in mine}
        h2d = sum(sum(host[c][k].numel() * host[c][k].element_size()
                      for k in ("upper_indptr", "upper_col", "upper_count", "weight")) for c in mine)
        h2d += sum(sum(t.numel() * t.element_size() for t in hwin[c]) for c in mine)
        d2h = acc.numel() * 8

        s_copy, s_comp = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ASYNC = _native.PUP_F_ASYNC
        # upload order: one small chromosome first (its upload is the only one nothing overlaps), then big to small
        e2e_order = sorted(mine, key=lambda c: -windows[c]["nb"])
        if len(e2e_order) > 2:
            e2e_order = [e2e_order[-1]] + e2e_order[:-1]
        # persistent device landing buffers for the window arrays (filled from the pinned host arrays every step)
        dwin_e2e = {c: tuple(torch.empty_like(t, device=dev) for t in hwin[c]) for c in mine}

        def e2e_step():
            # every upload (matrices via pup_region_create_upper, window arrays via pup_upload) travels on the
            # library's upload stream in consumption order and runs ahead of the kernels; chromosome k+1 is
            # mirrored / normalised / indexed on s_copy while chromosome k piles up on s_comp
            main_stream = torch.cuda.current_stream(dev)
            s_comp.wait_stream(main_stream)
            s_copy.wait_stream(main_stream)
            with torch.cuda.stream(s_comp):
                acc.zero_()

            def upload(c):
                h = host[c]
                reg = _native.Region(local_rank, windows[c]["nb"], h["upper_indptr"], h["upper_col"], h["upper_count"],
                                     h["weight"], None, None, ignore_diags=2, flags=ASYNC, stream=s_copy.cuda_stream,
                                     upper=True)
                for d, hsrc in zip(dwin_e2e[c], hwin[c]):
                    _native.upload(local_rank, d, hsrc, stream=s_copy.cuda_stream)
                ev = s_copy.record_event()  # matrix indexed and window arrays landed
                return reg, ev

            nxt = upload(e2e_order[0]) if e2e_order else None
            for k, c in enumerate(e2e_order):
                reg, ready = nxt
                nxt = upload(e2e_order[k + 1]) if k + 1 < len(e2e_order) else None
                s_comp.wait_event(ready)
                r0, c0, sl = dwin_e2e[c]
                reg.accumulate(r0, c0, sl, W, n_slots, flags | ASYNC, acc, stream=s_comp.cuda_stream)
                done = s_comp.record_event()
                s_copy.wait_event(done)
                with torch.cuda.stream(s_copy):
                    reg.close()
            main_stream.wait_stream(s_comp)
            main_stream.wait_stream(s_copy)
            if dist is not None:
                dist.all_reduce(acc)
            hacc.copy_(acc, non_blocking=True)
            torch.cuda.synchronize(dev)

        e2e_step()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        a1.record()
        barrier()
        ems = torch.tensor([a0.elapsed_time(a1) / args.e2e_steps], dtype=torch.float64, device=dev)
        bts = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            dist.all_reduce(bts)
        e2e = {"value": n_valid_total / (float(ems.item()) / 1e3), "unit": "pile-ups/s", "ms_per_step": float(ems.item()),
               "h2d_bytes_per_step": int(bts[0].item()), "d2h_bytes_per_step": int(bts[1].item()),
               "steps": args.e2e_steps,
               "what": "pup_region_create_upper + pup_accumulate per chromosome with pinned HOST buffers: the cooler-style "
                       "upper-triangle pixels (indptr, col, count), weights and window arrays are uploaded, mirrored / "
                       "normalised / indexed on the device and piled up; uploads run ahead on copy streams in consumption "
                       "order, chromosome k+1 is indexed while chromosome k piles up; D2H of the accumulators at the end"}

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
    # per rank: bytes of its windows / its main-kernel time; report the slowest rank's kernel (rank with max main ms)
    per_rank = [(float(s[0]), float(s[2]) / args.steps) for s in allstats]
    tot_bytes = sum(b for b, _ in per_rank)
    slow_b, slow_ms = max(per_rank, key=lambda x: x[1])
    achieved = (slow_b / 1e9) / (slow_ms / 1e3) if slow_ms > 0 else 0.0
    n_main = max(1, phases["main"][1])
    # DRAM traffic of the dominant kernel from the committed ncu capture (profiles/r1_k_pileup_main_ncu.md): the
    # chr1 launch moved 4.82 GB (read+write) for 16.6 GB of algorithmic bytes; scaled to the average launch here
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        traffic = int(tr["dram_bytes"] / tr["algorithmic_bytes"] * (slow_b / max(1, len(my_units))))
    except Exception:
        pass
    roofline = {
        "kernel": "k_pileup_main", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "traffic_source": "ncu --set full capture of the chr1 launch (profiles/r1_k_pileup_main_ncu.md), scaled by "
                          "algorithmic bytes to the average launch",
        "algorithmic_bytes_per_step": int(tot_bytes), "algorithmic_bytes_per_launch": int(slow_b / max(1, len(my_units))),
        "kernel_ms_per_step": slow_ms, "launches_per_step": n_main // args.steps,
        "avg_launch_ms": slow_ms / max(1, n_main // args.steps),
        "stored_pixels_in_windows_per_step": int(sum(float(s[1]) for s in allstats)),
        "phase_ms_per_step": {k: v[0] / args.steps for k, v in phases.items()},
    }

    # ---- CPU baseline: restated reference path on a bounded sample, 1 core
    cpu = None
    if not args.no_cpu and world == 1:
        sample_regions = [c for c in CPU_SAMPLE_REGIONS if c in mine and "col" in host[c]] or mine[:2]
        per = max(1, args.cpu_sample // len(sample_regions))
        rh = {c: dict(nb=windows[c]["nb"], **{k: host[c][k].numpy() for k in ("indptr", "col", "count", "weight")}) for c in sample_regions}
        cpu_setup(rh, windows, per)
        for c in sample_regions:
            _cpu_worker_prepare(c)
        t0 = time.perf_counter()
        nwin = sum(_cpu_worker_step(c) for c in sample_regions)
        dt = time.perf_counter() - t0
        cpu = {"value": nwin / dt, "unit": "pile-ups/s", "cores": 1, "kind": "port",
               "sample": f"{nwin} windows drawn uniformly from {','.join(sample_regions)} of the same workload "
                         f"({dt:.1f} s; matrix fetch+balancing untimed), oracle restatement of _stream_snips+_add_snip",
               "host_cpus": os.cpu_count()}

    line = {
        "metric": "pile-ups/sec (1e6 ROIs, 10kb bins, pad=41)", "value": value, "unit": "pile-ups/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(sum(float(s[4]) for s in allstats)),
        "roofline": roofline, "cpu_baseline": cpu,
        "windows_accumulated_per_step": n_valid_total,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_reference(args, names, sizes, windows, config, dev):
    """--impl reference: the restated reference CPU path on all usable host cores, bounded sample per step."""
    import torch

    from coolpuppy_b200.synthetic import synthetic_region

    regions_host = {}
    for ci, c in enumerate(names):
        t = synthetic_region(windows[c]["nb"], depth=args.depth, seed=1234 + ci, device=dev, nan_frac=0.03)
        regions_host[c] = dict(nb=t["nb"], **{k: t[k].cpu().numpy() for k in ("indptr", "col", "count", "weight")})
        del t
        torch.cuda.empty_cache()
    cpu_setup(regions_host, windows, args.ref_sample)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    value, sec_per_step, per_step, nproc = run_cpu_pool(names, cores, args.steps, args.warmup)
    sample = (f"{per_step} windows per step ({args.ref_sample} drawn uniformly per chromosome), one process per "
              f"chromosome task on {nproc} of {cores} cores like Pool.starmap over regions (coolpup.py:1502-1508); "
              "matrix fetch+balancing untimed")
    line = {
        "impl": "reference", "metric": "pile-ups/sec (1e6 ROIs, 10kb bins, pad=41)", "value": value, "unit": "pile-ups/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": "pile-ups/s", "cores": nproc, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pile-ups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

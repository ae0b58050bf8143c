"""Diagnostic: per-chromosome timeline of the end-to-end step (host buffers), run on the GPU box.

Prints, for every chromosome, when its upload+indexing finished on the copy stream and when its pile-up started /
finished on the compute stream (ms since the start of the step), to see whether uploads overlap pile-ups.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse

import torch

import bench
from coolpuppy_b200 import _native
from coolpuppy_b200.synthetic import synthetic_region

a = argparse.Namespace(chroms="all", pairs=1_000_000, nshifts=10, depth=500.0)
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
sizes = bench.chromsizes(a)
names = list(sizes)
windows, _, _ = bench.build_windows(a, sizes)
host, hwin = {}, {}
for ci, c in enumerate(names):
    t = synthetic_region(windows[c]["nb"], depth=a.depth, seed=1234 + ci, device=dev, nan_frac=0.03)
    host[c] = {k: t[k].cpu().pin_memory() for k in ("upper_indptr", "upper_col", "upper_count", "weight")}
    hwin[c] = tuple(torch.from_numpy(windows[c][k]).pin_memory() for k in ("r0", "c0", "slot"))
    del t
torch.cuda.empty_cache()
W, n_slots = bench.W, 2
acc = torch.zeros(n_slots * _native.acc_stride(W), dtype=torch.float64, device=dev)
s_copy, s_comp, s_up = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
UP = os.environ.get("E2E_WIN_UPLOAD", "1") == "1"  # window arrays uploaded ahead on their own stream (as bench.py does)
ASYNC = _native.PUP_F_ASYNC
order = sorted(names, key=lambda c: -windows[c]["nb"]) if os.environ.get("E2E_ORDER") == "big" else names
if os.environ.get("E2E_ONLY"):
    order = os.environ["E2E_ONLY"].split(",")


def step(record):
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t0 = ev()
    main = torch.cuda.current_stream(dev)
    t0.record(main)
    s_comp.wait_stream(main)
    s_copy.wait_stream(main)
    rows = []
    dwin, wev = {}, {}
    s_up.wait_stream(main)

    if os.environ.get("E2E_WIN_FIRST") == "1":  # experiment: every window array goes up before the first matrix
        with torch.cuda.stream(s_up):
            for c in order:
                dwin[c] = tuple(t.to(dev, non_blocking=True) for t in hwin[c])
                for t in dwin[c]:
                    t.record_stream(s_comp)
                wev[c] = s_up.record_event()

    def upload(c):
        h = host[c]
        if UP and c not in dwin:
            with torch.cuda.stream(s_up):
                dwin[c] = tuple(t.to(dev, non_blocking=True) for t in hwin[c])
                for t in dwin[c]:
                    t.record_stream(s_comp)
                wev[c] = s_up.record_event()
        with torch.cuda.stream(s_copy):
            reg = _native.Region(0, windows[c]["nb"], h["upper_indptr"], h["upper_col"], h["upper_count"], h["weight"], None,
                                 None, ignore_diags=2, flags=ASYNC, stream=s_copy.cuda_stream, upper=True)
            e = ev()
            e.record(s_copy)
        return reg, e

    import time as _t

    h_up = h_acc = 0.0
    th = _t.perf_counter()
    nxt = upload(order[0])
    h_up += _t.perf_counter() - th
    for k, c in enumerate(order):
        reg, ready = nxt
        th = _t.perf_counter()
        nxt = upload(order[k + 1]) if k + 1 < len(order) else None
        h_up += _t.perf_counter() - th
        s_comp.wait_event(ready)
        a0, a1 = ev(), ev()
        a0.record(s_comp)
        r0, c0, sl = hwin[c]
        if UP:
            s_comp.wait_event(wev[c])
            r0, c0, sl = dwin[c]
        th = _t.perf_counter()
        reg.accumulate(r0, c0, sl, W, n_slots, ASYNC, acc, stream=s_comp.cuda_stream)
        h_acc += _t.perf_counter() - th
        a1.record(s_comp)
        s_copy.wait_event(a1)
        with torch.cuda.stream(s_copy):
            reg.close()
        rows.append((c, ready, a0, a1))
    if record:
        print(f"host time inside Region() calls {h_up * 1e3:.1f} ms, inside accumulate() calls {h_acc * 1e3:.1f} ms")
    main.wait_stream(s_comp)
    main.wait_stream(s_copy)
    main.wait_stream(s_up)
    t1 = ev()
    t1.record(main)
    torch.cuda.synchronize()
    if record:
        for c, ready, a0, a1 in rows:
            print(f"{c:6s} upload+index done {t0.elapsed_time(ready):7.1f}   pile-up {t0.elapsed_time(a0):7.1f} -> {t0.elapsed_time(a1):7.1f}")
    return t0.elapsed_time(t1)


import time

step(False)
step(False)
_native.timing_enable(True)
_native.timing_read(reset=True)
t0 = time.perf_counter()
print("step ms:", step(True), "host wall ms:", (time.perf_counter() - t0) * 1e3)
print("phases (ms, spans):", _native.timing_read(reset=True))
_native.timing_enable(False)
print("step ms:", step(False))

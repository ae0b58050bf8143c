"""Device-generated windows vs the host builder on the bench workload (GPU box): exact array comparison per region."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from coolpuppy_b200 import _native as nat
from coolpuppy_b200 import coolpup as cp
from coolpuppy_b200.coolio import ChromCooler

args = bench.parse_args()
wl = bench.WORKLOADS[args.workload]
sizes = bench.chromsizes(args)
features, n_pairs = bench.make_features(args, sizes)
nbins = sum(-(-L // bench.BINSIZE) for L in sizes.values())
clr = ChromCooler(sizes, bench.BINSIZE, {}, {"weight": np.ones(nbins)})
np.random.seed(0)
cc = cp.CoordCreator(features, bench.BINSIZE, features_format="bed", flank=wl["flank"], nshifts=wl["nshifts"], mindist="auto", seed=0)
pu = cp.PileUpper(clr, cc, clr_weight_name="weight", control=wl["nshifts"] > 0)
plan = pu._plan([], False, None, None)
plan["band_edges"] = None
state0 = np.random.get_state()
host = pu._prepare(plan, None, None)
np.random.set_state(state0)
job = pu._prepare_device(plan, None, None)
dev = torch.device("cuda", 0)
rng = nat.DeviceRng(0)
by_name = {b["name"]: b for b in host["built"]}
bad = 0
for it in job["items"]():
    dbin = None
    if len(it["segs"]):
        dbin = torch.empty(int(it["total"]) * it["nctrl"], dtype=torch.int32, device=dev)
        rng.control_shifts(it["segs"], pu.CC.minshift, pu.CC.maxshift, pu.resolution, dbin)
    if not it["owned"]:
        continue
    b = by_name[it["name"]]
    n_all = int(it["total"]) * (1 + it["nctrl"])
    outs = tuple(torch.full((n_all,), -7, dtype=torch.int32, device=dev) for _ in range(3))
    nat.pair_windows_device(0, it["stbin"], it["center"], pu.CC.mindist, pu.CC.maxdist, it["nctrl"], it["per_offset"], dbin,
                            it["nb"], job["W"], None, None, None, 0, 0, False, None, None, job["nk"], job["nf"], 0, len(it["center"]),
                            None, it["index"], outs[0], outs[1], outs[2])
    torch.cuda.synchronize()
    for nm, got, want in zip(("r0", "c0", "slot"), outs, (b["w_r0"], b["w_c0"], b["slot"])):
        g = got.cpu().numpy().astype(np.int64)
        w = np.asarray(want).astype(np.int64)
        if not np.array_equal(g, w):
            d = np.nonzero(g != w)[0]
            bad += 1
            print(it["name"], nm, "differs at", len(d), "of", len(w), "first", d[:5], g[d[:5]], w[d[:5]], "m", len(it["center"]))
            # which offsets?
            base = np.concatenate([[0], np.cumsum(it["per_offset"])])
            blk = np.searchsorted(base * (1 + it["nctrl"]), d[:5], side="right") - 1
            print("   offsets", blk, "q", it["per_offset"][blk], "block start", (base * (1 + it["nctrl"]))[blk])
print("regions with differences:", bad)
W = job["W"]
tot = 0
for b in host["built"]:
    tot += int(np.count_nonzero(b["valid"]))
print("host-builder valid windows:", tot, "of", sum(len(b["valid"]) for b in host["built"]))

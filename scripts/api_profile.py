"""cProfile of coolpuppy_b200.coolpup.pileup() on a bench workload (GPU box): where does the host time go?"""
import cProfile
import os
import pstats
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from coolpuppy_b200 import coolpup as cp

sys.argv = [sys.argv[0]] + sys.argv[1:]
args = bench.parse_args()
wl = bench.WORKLOADS[args.workload]
dev = torch.device("cuda", 0)
sizes = bench.chromsizes(args)
features, n_pairs = bench.make_features(args, sizes)
dev_data, host_data = bench.generate_genome(args, sizes, dev, keep_device=set(sizes) if wl["expected"] else set(), keep_host=set(sizes), pin=True)
expected_df = bench.expected_table(dev_data) if wl["expected"] else None
clr = bench.host_cooler(sizes, host_data, pin=False)
kw = dict(wl["kwargs"], flank=wl["flank"], clr_weight_name="weight", device=0)
if wl["expected"]:
    kw["expected_df"] = expected_df
import logging
logging.getLogger("coolpuppy").setLevel(logging.WARNING)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    cp.pileup(clr, features, **kw)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        cp.pileup(clr, features, **kw)
        torch.cuda.synchronize()
    pr.disable()
print(cp._LAST_STATS)
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

"""CUDA-event time of the MT19937 control-shift replay (k_mt_shifts) for a configs[3]-sized stream of draws."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from coolpuppy_b200 import _native

dev = torch.device("cuda", 0)
_native.require_device()
np.random.seed(0)
rng = np.random.default_rng(1)
for total, nseg in ((10_000_000, 6000), (10_000_000, 24), (1_000_000, 600)):
    seg = rng.multinomial(total, np.ones(nseg) / nseg).astype(np.int64)
    dbin = torch.empty(total, dtype=torch.int32, device=dev)
    r = _native.DeviceRng(0)
    r.control_shifts(seg[:10], 100000, 1000000, 10000, dbin)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r.control_shifts(seg, 100000, 1000000, 10000, dbin)
    b.record()
    torch.cuda.synchronize()
    print(total, nseg, "ms", a.elapsed_time(b))
    r.close()

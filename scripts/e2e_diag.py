"""Diagnostic: where does the end-to-end (host buffers) time go?  Run on the GPU box."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from coolpuppy_b200 import _native
from coolpuppy_b200.synthetic import synthetic_region, HG38

dev = torch.device("cuda", 0)
names = list(HG38)
host = {}
for ci, c in enumerate(names):
    nb = -(-HG38[c] // 10000)
    t = synthetic_region(nb, depth=500.0, seed=1234 + ci, device=dev, nan_frac=0.03)
    host[c] = dict(nb=nb, **{k: t[k].cpu().pin_memory() for k in ("upper_indptr", "upper_col", "upper_count", "weight")})
    del t
torch.cuda.empty_cache()
st = torch.cuda.current_stream(dev).cuda_stream
total = sum(sum(v.numel() * v.element_size() for k, v in h.items() if k != "nb") for h in host.values())

def upload_all(flags):
    regs = []
    for c in names:
        h = host[c]
        regs.append(_native.Region(0, h["nb"], h["upper_indptr"], h["upper_col"], h["upper_count"], h["weight"], None, None,
                                   ignore_diags=2, flags=flags, stream=st, upper=True))
    torch.cuda.synchronize()
    for r in regs:
        r.close()

def h2d_only():
    bufs = []
    for c in names:
        h = host[c]
        bufs.append([h[k].to(dev, non_blocking=True) for k in ("upper_indptr", "upper_col", "upper_count", "weight")])
    torch.cuda.synchronize()

for name, f in (("H2D only (torch copies)", h2d_only), ("create_upper x24 async", lambda: upload_all(_native.PUP_F_ASYNC)),
                ("create_upper x24 sync", lambda: upload_all(0))):
    f()
    t0 = time.perf_counter()
    for _ in range(3):
        f()
    print(f"{name}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms  ({total / 1e9:.2f} GB)")

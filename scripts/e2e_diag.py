"""Diagnostic: where does the end-to-end (host buffers) time go?  Run on the GPU box."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from coolpuppy_b200 import _native
from coolpuppy_b200.synthetic import synthetic_region

dev = torch.device("cuda", 0)
nb = 24896
t = synthetic_region(nb, depth=500.0, seed=1234, device=dev, nan_frac=0.03)
host = {k: t[k].cpu().pin_memory() for k in ("upper_indptr", "upper_col", "upper_count", "weight", "indptr", "col", "count")}
nbytes_u = sum(host[k].numel() * host[k].element_size() for k in ("upper_indptr", "upper_col", "upper_count", "weight"))
nbytes_s = sum(host[k].numel() * host[k].element_size() for k in ("indptr", "col", "count", "weight"))
st = torch.cuda.current_stream(dev).cuda_stream

def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

# raw H2D
buf = torch.empty_like(t["upper_col"])
ms = timeit(lambda: buf.copy_(host["upper_col"], non_blocking=True))
print(f"H2D pinned {host['upper_col'].numel()*4/1e6:.0f} MB: {ms:.2f} ms = {host['upper_col'].numel()*4/ms/1e6:.1f} GB/s")

def mk_upper(flags=0):
    r = _native.Region(0, nb, host["upper_indptr"], host["upper_col"], host["upper_count"], host["weight"], None, None,
                       ignore_diags=2, flags=flags, stream=st, upper=True)
    r.close()
def mk_sym(flags=0):
    r = _native.Region(0, nb, host["indptr"], host["col"], host["count"], host["weight"], None, None, ignore_diags=2, flags=flags, stream=st)
    r.close()
def mk_upper_dev():
    r = _native.Region(0, nb, t["upper_indptr"], t["upper_col"], t["upper_count"], t["weight"], None, None, ignore_diags=2, stream=st, upper=True)
    r.close()
def mk_sym_dev():
    r = _native.Region(0, nb, t["indptr"], t["col"], t["count"], t["weight"], None, None, ignore_diags=2, stream=st)
    r.close()
print(f"create_upper host ({nbytes_u/1e6:.0f} MB): {timeit(mk_upper):.2f} ms   async: {timeit(lambda: mk_upper(_native.PUP_F_ASYNC)):.2f} ms")
print(f"create_sym   host ({nbytes_s/1e6:.0f} MB): {timeit(mk_sym):.2f} ms   async: {timeit(lambda: mk_sym(_native.PUP_F_ASYNC)):.2f} ms")
print(f"create_upper device-resident inputs: {timeit(mk_upper_dev):.2f} ms ; create_sym device-resident: {timeit(mk_sym_dev):.2f} ms")

"""ctypes binding of ``libpileup_b200.so`` (C ABI: ``include/pileup_b200.h``).

There is deliberately no fallback: if the library is missing, or no CUDA device
is visible, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PUP_F_OOE = 1
PUP_F_EXPCTRL = 2
PUP_F_COVERAGE = 4
PUP_F_NODIAG = 8
PUP_F_ASYNC = 16
PUP_F_LOCAL = 32

_LIB = None
_PATH = os.environ.get("PUP_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpileup_b200.so")  # PUP_LIB: tuning builds

SYMBOLS = [
    "pup_abi_version", "pup_last_error", "pup_device_count", "pup_region_create", "pup_region_create_upper",
    "pup_region_destroy", "pup_upload", "pup_expected_cis", "pup_pair_windows_count", "pup_pair_windows_fill",
    "pup_region_device_bytes", "pup_acc_stride", "pup_accumulate", "pup_accumulate_region", "pup_acc_export",
    "pup_last_launches", "pup_algorithmic_bytes", "pup_timing_enable", "pup_timing_read", "pup_stripes",
    "pup_pair_windows_count_range", "pup_accumulate_rescaled", "pup_rng_create", "pup_rng_read", "pup_rng_destroy", "pup_control_shifts", "pup_pair_windows_device",
]


class NativeError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_PATH):
        raise NativeError(
            f"{_PATH} is missing: build it with `python -m coolpuppy_b200.build` "
            "(there is no CPU fallback for the pile-up path)"
        )
    L = C.CDLL(_PATH)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint
    L.pup_abi_version.restype = C.c_int
    L.pup_last_error.restype = C.c_char_p
    L.pup_last_launches.restype = C.c_int
    L.pup_device_count.argtypes = [C.POINTER(C.c_int)]
    L.pup_region_create.argtypes = [C.c_int, i32, i64, vp, vp, vp, vp, vp, vp, C.c_int, u32, vp, C.POINTER(vp)]
    L.pup_region_create_upper.argtypes = L.pup_region_create.argtypes
    L.pup_region_destroy.argtypes = [vp]
    L.pup_upload.argtypes = [C.c_int, vp, vp, i64, vp]
    L.pup_expected_cis.argtypes = [C.c_int, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.pup_pair_windows_count.argtypes = [i32, vp, C.c_double, C.c_double, vp]
    L.pup_pair_windows_count.restype = i64
    L.pup_pair_windows_count_range.argtypes = [i32, vp, C.c_double, C.c_double, i32, i32, vp]
    L.pup_pair_windows_count_range.restype = i64
    L.pup_pair_windows_fill.argtypes = [i32, vp, vp, C.c_double, C.c_double, i32, vp, vp, vp, vp, vp, vp, vp]
    L.pup_region_device_bytes.argtypes = [vp]
    L.pup_region_device_bytes.restype = i64
    L.pup_acc_stride.argtypes = [C.c_int]
    L.pup_acc_stride.restype = i64
    L.pup_accumulate.argtypes = [vp, i64, vp, vp, vp, C.c_int, C.c_int, u32, vp, vp, C.POINTER(i64)]
    L.pup_accumulate_rescaled.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, u32, vp, vp, C.POINTER(i64)]
    L.pup_accumulate_region.argtypes = [C.c_int, i32, i64, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, C.c_int, C.c_int,
                                        C.c_int, u32, vp, vp, C.POINTER(i64)]
    L.pup_acc_export.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.pup_algorithmic_bytes.argtypes = [vp, i64, vp, vp, C.c_int, u32, vp, C.POINTER(i64), C.POINTER(i64)]
    L.pup_stripes.argtypes = [vp, i64, vp, vp, C.c_int, vp, vp, vp]
    L.pup_rng_create.argtypes = [C.c_int, vp, C.c_int, vp, C.POINTER(vp)]
    L.pup_rng_read.argtypes = [vp, vp, C.POINTER(C.c_int), vp]
    L.pup_rng_destroy.argtypes = [vp]
    L.pup_control_shifts.argtypes = [vp, i64, vp, i64, i64, C.c_double, vp, vp]
    L.pup_pair_windows_device.argtypes = [C.c_int, i32, vp, vp, C.c_double, C.c_double, i32, vp, vp, i32, C.c_int, vp,
                                          vp, vp, i32, i64, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, i32, i32, vp, i32,
                                          vp, vp, vp, vp, vp, vp]
    L.pup_timing_enable.argtypes = [C.c_int]
    L.pup_timing_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise NativeError(f"libpileup_b200 error {rc}: {lib().pup_last_error().decode()}")


def device_count():
    n = C.c_int(0)
    rc = lib().pup_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def require_device():
    if device_count() < 1:
        raise NativeError("no CUDA device visible: the pile-up path has no CPU fallback")


def ptr(a, dtype=None):
    """Raw address of a numpy array / torch tensor (host or device) or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if dtype is not None and a.dtype != np.dtype(dtype):
            raise TypeError(f"expected {np.dtype(dtype)}, got {a.dtype}")
        if not a.flags.c_contiguous:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError(f"unsupported buffer type {type(a)}")


def alloc_accumulator(n_doubles, device):
    """Zeroed fp64 accumulator in HBM (a torch tensor, so that torch.distributed can all-reduce it)."""
    import torch

    if not torch.cuda.is_available():
        raise NativeError("no CUDA device visible: the pile-up path has no CPU fallback")
    return torch.zeros(int(n_doubles), dtype=torch.float64, device=torch.device("cuda", device))


def current_stream(device):
    """The caller's current CUDA stream on ``device`` as a raw ``cudaStream_t`` value."""
    import torch

    return torch.cuda.current_stream(torch.device("cuda", device)).cuda_stream


def acc_stride(W):
    return int(lib().pup_acc_stride(int(W)))


def acc_counts(acc, W, n_slots):
    """Windows accumulated per slot (the ``n`` field of every slot of a torch accumulator) as a host int64 array."""
    stride = acc_stride(W)
    n = acc.view(int(n_slots), stride)[:, stride - 8]  # AccLayout: off_n = stride - 8
    return np.rint(n.cpu().numpy()).astype(np.int64)


def device_windows_supported():
    """True when windows can be generated on the device (always, with the real library; the CPU test emulator
    replaces this by False)."""
    return True


def acc_has_inf(acc, W, n_slots):
    """True when some pile-up sum of a torch accumulator is infinite (x / 0 pixels poison their cell)."""
    import torch

    stride = acc_stride(W)
    return bool(torch.isinf(acc.view(int(n_slots), stride)[:, : int(W) * int(W)]).any().item())


def make_pipeline(device, W, n_slots, flags):
    """The two-stream region pipeline (:mod:`coolpuppy_b200.pipeline`) on ``device``."""
    from .pipeline import RegionPipeline

    return RegionPipeline(device, W, n_slots, flags)


class Region:
    """A region matrix resident in HBM (``pup_region_t``)."""

    def __init__(self, device, nb, indptr, col, count, weight=None, expected=None, coverage=None, ignore_diags=2,
                 flags=0, stream=0, upper=False):
        """``upper=True``: ``indptr/col/count`` hold the upper triangle as cooler stores it (columns >= nb are
        dropped; the lower triangle is mirrored in on the device only when the diagonal mask keeps it, i.e.
        ``ignore_diags < 0``); else the symmetric-filled CSR."""
        self._h = C.c_void_p()
        self.nb = int(nb)
        self.nnz = int(col.shape[0]) if col is not None else 0
        self.device = device
        self.balanced = weight is not None
        fn = lib().pup_region_create_upper if upper else lib().pup_region_create
        check(fn(device, self.nb, self.nnz, ptr(indptr), ptr(col), ptr(count), ptr(weight), ptr(expected),
                 ptr(coverage), int(ignore_diags), int(flags) & (PUP_F_OOE | PUP_F_NODIAG | PUP_F_ASYNC), stream,
                 C.byref(self._h)))

    @property
    def device_bytes(self):
        return int(lib().pup_region_device_bytes(self._h))

    def accumulate(self, r0, c0, slot, W, n_slots, flags, acc, stream=0, want_n_valid=False):
        n = int(r0.shape[0])
        nv = C.c_int64(0)
        check(lib().pup_accumulate(self._h, n, ptr(r0), ptr(c0), ptr(slot), int(W), int(n_slots),
                                   int(flags) & (PUP_F_EXPCTRL | PUP_F_COVERAGE | PUP_F_ASYNC), ptr(acc), stream,
                                   C.byref(nv) if want_n_valid else None))
        return nv.value if want_n_valid else None

    def accumulate_rescaled(self, r0, c0, h, w, slot, mode, rescale_size, n_slots, flags, acc, stream=0,
                            want_n_valid=False):
        """``pup_accumulate_rescaled``: windows ``[r0, r0 + h) x [c0, c0 + w)`` zoomed to ``rescale_size`` squared."""
        n = int(r0.shape[0])
        nv = C.c_int64(0)
        check(lib().pup_accumulate_rescaled(self._h, n, ptr(r0), ptr(c0), ptr(h), ptr(w), ptr(slot), ptr(mode),
                                            int(rescale_size), int(n_slots),
                                            int(flags) & (PUP_F_COVERAGE | PUP_F_LOCAL | PUP_F_ASYNC), ptr(acc), stream,
                                            C.byref(nv) if want_n_valid else None))
        return nv.value if want_n_valid else None

    def stripes(self, r0, c0, W, stream=0):
        """(horizontal, vertical) centre stripes of every window as host arrays [n, W]."""
        n = int(r0.shape[0])
        hor = np.empty((n, int(W)), dtype=np.float64)
        ver = np.empty((n, int(W)), dtype=np.float64)
        check(lib().pup_stripes(self._h, n, ptr(r0), ptr(c0), int(W), ptr(hor), ptr(ver), stream))
        return hor, ver

    def algorithmic_bytes(self, r0, c0, W, flags=0, stream=0):
        b, z = C.c_int64(0), C.c_int64(0)
        check(lib().pup_algorithmic_bytes(self._h, int(r0.shape[0]), ptr(r0), ptr(c0), int(W), int(flags), stream,
                                          C.byref(b), C.byref(z)))
        return b.value, z.value

    def close(self):
        if self._h:
            lib().pup_region_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def accumulate_region(device, nb, indptr, col, count, weight, expected, coverage, r0, c0, slot, W, ignore_diags,
                      n_slots, flags, acc, stream=0):
    """One-shot ``pup_accumulate_region`` (everything may be host memory). Returns n_valid."""
    nv = C.c_int64(0)
    check(lib().pup_accumulate_region(device, int(nb), int(col.shape[0]), ptr(indptr), ptr(col), ptr(count),
                                      ptr(weight), ptr(expected), ptr(coverage), int(r0.shape[0]), ptr(r0), ptr(c0),
                                      ptr(slot), int(W), int(ignore_diags), int(n_slots), int(flags), ptr(acc),
                                      stream, C.byref(nv)))
    return nv.value


def upload(device, dst, src, stream=0):
    """``pup_upload``: async H2D copy of ``src`` (host tensor / array, pinned for real asynchrony) into the device
    tensor ``dst`` through the library's upload stream (FIFO with the region matrices); ``stream`` waits for it."""
    nbytes = src.numel() * src.element_size() if hasattr(src, "numel") else src.nbytes
    check(lib().pup_upload(device, ptr(dst), ptr(src), int(nbytes), stream))


def pair_windows_count(center, mindist, maxdist, k_lo=0, k_hi=None):
    """``pup_pair_windows_count[_range]`` (host code): kept all-vs-all pairs per offset, int64[m] (only the pairs
    whose first feature lies in ``[k_lo, k_hi)`` when given)."""
    center = np.ascontiguousarray(center, dtype=np.float64)
    q = np.zeros(max(1, center.shape[0]), dtype=np.int64)
    m = int(center.shape[0])
    total = lib().pup_pair_windows_count_range(m, ptr(center), float(mindist), float(maxdist), int(k_lo),
                                               m if k_hi is None else int(k_hi), ptr(q))
    if total < 0:
        raise NativeError(f"libpileup_b200: {lib().pup_last_error().decode()}")
    return q[: center.shape[0]], int(total)


def pair_windows_fill(stbin, center, mindist, maxdist, nctrl, dbin, total):
    """``pup_pair_windows_fill`` (host code): ``(st1, st2, kind, idx1, idx2, distance)`` of the region's windows."""
    stbin = np.ascontiguousarray(stbin, dtype=np.int64)
    center = np.ascontiguousarray(center, dtype=np.float64)
    n = int(total) * (1 + int(nctrl))
    st1, st2 = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
    kind = np.empty(n, dtype=np.int8)
    i1, i2 = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
    dist = np.empty(n, dtype=np.float64)
    if dbin is not None:
        dbin = np.ascontiguousarray(dbin, dtype=np.int64)
    check(lib().pup_pair_windows_fill(int(center.shape[0]), ptr(stbin), ptr(center), float(mindist), float(maxdist),
                                      int(nctrl), ptr(dbin), ptr(st1), ptr(st2), ptr(kind), ptr(i1), ptr(i2), ptr(dist)))
    return st1, st2, kind, i1, i2, dist


class DeviceRng:
    """numpy's global legacy MT19937 state on the device (``pup_rng_t``): loaded from ``np.random.get_state()``,
    advanced by :meth:`control_shifts`, written back with :meth:`store` so that host code continues the stream."""

    def __init__(self, device, stream=0):
        st = np.random.get_state()
        if st[0] != "MT19937":
            raise NativeError("np.random is not the legacy MT19937 generator")
        self.device = int(device)
        self._legacy = st
        self._h = C.c_void_p()
        key = np.ascontiguousarray(st[1], dtype=np.uint32)
        check(lib().pup_rng_create(self.device, ptr(key), int(st[2]), stream, C.byref(self._h)))

    def control_shifts(self, seg_sizes, minshift, maxshift, resolution, dbin, stream=0):
        """Replay ``randint(minshift, maxshift, n); choice([-1, 1], n)`` for every ``n`` of ``seg_sizes``; ``dbin``:
        int32 device tensor of ``sum(seg_sizes)`` bins (``None``: only advance the stream)."""
        seg = np.ascontiguousarray(seg_sizes, dtype=np.int64)
        if seg.size == 0:
            return
        check(lib().pup_control_shifts(self._h, int(seg.size), ptr(seg), int(minshift), int(maxshift), float(resolution),
                                       ptr(dbin), stream))

    def store(self, stream=0):
        """Write the advanced state back into ``np.random`` (synchronises ``stream``)."""
        key = np.empty(624, dtype=np.uint32)
        pos = C.c_int(0)
        check(lib().pup_rng_read(self._h, ptr(key), C.byref(pos), stream))
        st = self._legacy
        np.random.set_state((st[0], key, int(pos.value), st[3], st[4]))

    def close(self):
        if self._h:
            lib().pup_rng_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pair_windows_device(device, stbin, center, mindist, maxdist, nctrl, per_offset, dbin, nb, W, key1, key2, band_edges,
                        band_weight, flip_mode, swap_on_flip, flipval, ident, nk, nf, k_lo, k_hi, per_offset_part,
                        region_index, r0, c0, slot, first_seen=None, n_roi=None, stream=0):
    """``pup_pair_windows_device``: all-vs-all windows + slots of one region written into the device tensors
    ``r0 / c0 / slot`` in the reference's emission order."""
    m = int(center.shape[0])
    if int(r0.shape[0]) == 0:  # this part holds no window
        return
    check(lib().pup_pair_windows_device(
        int(device), m, ptr(stbin, np.int32), ptr(center, np.float64), float(mindist), float(maxdist),
        int(nctrl), ptr(per_offset, np.int64), ptr(dbin), int(nb), int(W), ptr(key1), ptr(key2), ptr(band_edges),
        0 if band_edges is None else int(band_edges.shape[0]), int(band_weight), int(flip_mode), int(bool(swap_on_flip)),
        ptr(flipval), ptr(ident), int(nk), int(nf), int(k_lo), int(k_hi), ptr(per_offset_part, np.int64),
        int(region_index), ptr(r0), ptr(c0), ptr(slot), ptr(first_seen), ptr(n_roi), stream))


def expected_cis_sums(device, nb, indptr_upper, col_upper, count_upper, weight=None, stream=0):
    """``pup_expected_cis``: per-diagonal ``(count_sum f64[nb], balanced_sum f64[nb] | None, n_valid i64[nb])`` of one
    region's upper triangle (host arrays out)."""
    nb = int(nb)
    cs = np.empty(nb, dtype=np.float64)
    bs = np.empty(nb, dtype=np.float64) if weight is not None else None
    nv = np.empty(nb, dtype=np.int64)
    check(lib().pup_expected_cis(device, nb, int(col_upper.shape[0]), ptr(indptr_upper), ptr(col_upper), ptr(count_upper),
                                 ptr(weight), ptr(cs), ptr(bs), ptr(nv), stream))
    return cs, bs, nv


def timing_enable(on=True):
    check(lib().pup_timing_enable(1 if on else 0))


def timing_read(reset=True):
    """{phase: (milliseconds, spans)} for phases 'plan', 'vector', 'main' (sparse pile-up kernel), 'dense_num',
    'dense_band' (dense pile-up kernel) since the last reset."""
    ms = (C.c_double * 5)()
    cnt = (C.c_int * 5)()
    check(lib().pup_timing_read(ms, cnt, 1 if reset else 0))
    return {k: (ms[i], cnt[i]) for i, k in enumerate(("plan", "vector", "main", "dense_num", "dense_band"))}


def acc_export(acc, W, n_slots, device=0, stream=0, want_expected=False, want_cov=False):
    """Decode an accumulator buffer into dict(sum, num, n[, cov_start, cov_end][, exp_sum, exp_num])."""
    W, n_slots = int(W), int(n_slots)
    out = {
        "sum": np.empty((n_slots, W, W), dtype=np.float64),
        "num": np.empty((n_slots, W, W), dtype=np.int64),
        "n": np.empty(n_slots, dtype=np.int64),
    }
    if want_cov:
        out["cov_start"] = np.empty((n_slots, W), dtype=np.float64)
        out["cov_end"] = np.empty((n_slots, W), dtype=np.float64)
    if want_expected:
        out["exp_sum"] = np.empty((n_slots, W, W), dtype=np.float64)
        out["exp_num"] = np.empty((n_slots, W, W), dtype=np.int64)
    check(lib().pup_acc_export(ptr(acc), W, n_slots, device, stream, ptr(out["sum"]), ptr(out["num"]), ptr(out["n"]),
                               ptr(out.get("cov_start")), ptr(out.get("cov_end")), ptr(out.get("exp_sum")),
                               ptr(out.get("exp_num"))))
    return out

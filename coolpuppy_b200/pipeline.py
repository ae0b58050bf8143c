"""Two-stream region pipeline: view region k+1 is uploaded and indexed while region k piles up.

The reference maps ``pileup_region`` over the view regions one after the other (or over ``Pool`` workers,
``coolpup.py:1495-1510``); each call reads its chromosome from the cooler, builds the CSR and walks the windows.
Here the same per-region work is split over two CUDA streams so that nothing waits for the host or for PCIe:

* *prepare* stream: ``pup_region_create[_upper](PUP_F_ASYNC)`` -- host->device copies of the region's pixels
  (through the library's upload FIFO), normalisation, strip layout, bucket table -- and the region's window arrays
  (``pup_upload`` of host arrays, or generated on the device);
* *compute* stream: ``pup_accumulate(PUP_F_ASYNC)`` of the previous region, which waits for that region's event.

No call blocks the host: the host is free to read / decompress / lay out the next region while the GPU works.
Host buffers handed to :meth:`submit` are kept alive until their copies have completed.
"""
from __future__ import annotations

from collections import deque

import numpy as np

from . import _native


_STREAMS = {}
_SCRATCH = {}  # (device, tag, parity) -> persistent int32 device tensor


def device_streams(device):
    """(prepare, compute, side) streams of ``device``, created once per process: torch's caching allocator keeps one
    memory pool per stream, so fresh streams per run would turn every scratch allocation into a cudaMalloc."""
    import torch

    if device not in _STREAMS:
        dev = torch.device("cuda", device)
        _STREAMS[device] = tuple(torch.cuda.Stream(dev) for _ in range(3))
    return _STREAMS[device]


class RegionPipeline:
    def __init__(self, device, W, n_slots, flags):
        import torch

        self.torch = torch
        self.device = int(device)
        self.dev = torch.device("cuda", self.device)
        self.W, self.n_slots, self.flags = int(W), int(n_slots), int(flags)
        self.main = torch.cuda.current_stream(self.dev)
        self.s_prep, self.s_comp, self.s_side = device_streams(self.device)
        self.s_side.wait_stream(self.main)
        self.s_prep.wait_stream(self.main)
        self.s_comp.wait_stream(self.main)
        self._pending = None
        self._submitted = 0
        self._keep = deque()  # (event, python objects whose memory an enqueued copy still reads)
        self.launches = 0
        self.regions = 0

    # -- helpers ----------------------------------------------------------------------------------
    def scratch_i32(self, tag, n, count=1, stream=None):
        """``count`` int32 device vectors of length ``n`` for the region being submitted, carved out of a persistent
        buffer per (tag, region parity).  Region k's buffers are reused by region k + 2, whose prepare-stream work is
        ordered after region k's pile-up (the ``wait_event(done)`` of :meth:`_compute`), so no allocator round trip
        and no ``record_stream`` bookkeeping is needed per region."""
        torch = self.torch
        key = (self.device, tag, self._submitted & 1)
        n = int(n)
        step = (n + 63) // 64 * 64
        buf = _SCRATCH.get(key)
        if buf is None or buf.numel() < count * step:
            with torch.cuda.stream(stream or self.s_prep):
                buf = torch.empty(max(count * step, 1024) * 5 // 4, dtype=torch.int32, device=self.dev)
            _SCRATCH[key] = buf
        return tuple(buf[i * step : i * step + n] for i in range(count))

    def _reap(self, everything=False):
        while self._keep and (everything or self._keep[0][0].query()):
            self._keep.popleft()

    def upload_windows(self, arrays):
        """Host int32 arrays -> device tensors through the library's upload FIFO (ordered with the matrix copies)."""
        torch = self.torch
        arrays = [np.ascontiguousarray(a, dtype=np.int32) for a in arrays]
        out = self.scratch_i32("windows", arrays[0].shape[0], count=len(arrays))
        with torch.cuda.stream(self.s_prep):
            for a, d in zip(arrays, out):
                if a.shape[0]:
                    _native.upload(self.device, d, a, stream=self.s_prep.cuda_stream)
                self._hold.append(a)
        return out

    # -- the pipeline -----------------------------------------------------------------------------
    def submit(self, region_kwargs, windows, acc, after=None, windows_on_device=None):
        """Enqueue one view region.

        ``region_kwargs``: arguments of :class:`_native.Region` (host or device arrays); ``windows``: ``(r0, c0,
        slot)`` host int32 arrays, or ``None`` with ``windows_on_device(stream) -> (r0, c0, slot)`` device tensors
        produced on the prepare stream; ``acc``: the device accumulator this region adds into;
        ``after(region, stream)``: optional extra work on the compute stream right after the pile-up (stripes).
        """
        torch = self.torch
        self._reap()
        # small per-bin host vectors go through pinned staging tensors: a pageable source would make the upload call
        # wait until every copy queued before it (the region's pixels, 100s of MB) has drained
        for k in ("weight", "expected", "coverage"):
            v = region_kwargs.get(k)
            if isinstance(v, np.ndarray):
                t = torch.empty(v.shape[0], dtype=torch.float64, pin_memory=True)
                t.numpy()[:] = v
                region_kwargs[k] = t
        self._hold = [v for v in region_kwargs.values() if isinstance(v, np.ndarray) or hasattr(v, "data_ptr")]
        with torch.cuda.stream(self.s_prep):
            region = _native.Region(self.device, flags=region_kwargs.pop("flags", 0) | _native.PUP_F_ASYNC,
                                    stream=self.s_prep.cuda_stream, **region_kwargs)
            self.launches += int(_native.lib().pup_last_launches())
            if windows is not None:
                wins = self.upload_windows(windows)
            else:
                wins = windows_on_device(self.s_prep)
            ready = self.s_prep.record_event()
        self._keep.append((ready, self._hold))
        self._hold = []
        self._submitted += 1
        prev, self._pending = self._pending, (region, wins, ready, acc, after)
        if prev is not None:
            self._compute(prev)
        return ready

    def _compute(self, item):
        region, wins, ready, acc, after = item
        torch = self.torch
        self.s_comp.wait_event(ready)
        r0, c0, slot = wins[:3]
        if r0.shape[0] and len(wins) >= 5:  # rescaled pile-up: windows of their own sizes, zoomed to W x W
            mode = wins[5] if len(wins) == 6 else None  # only with bare expected blocks as control snippets
            region.accumulate_rescaled(r0, c0, wins[3], wins[4], slot, mode, self.W, self.n_slots,
                                       self.flags | _native.PUP_F_ASYNC, acc, stream=self.s_comp.cuda_stream)
            self.launches += int(_native.lib().pup_last_launches())
        elif r0.shape[0]:
            region.accumulate(r0, c0, slot, self.W, self.n_slots, self.flags | _native.PUP_F_ASYNC, acc,
                              stream=self.s_comp.cuda_stream)
            self.launches += int(_native.lib().pup_last_launches())
        if after is not None:
            with torch.cuda.stream(self.s_comp):
                after(region, self.s_comp.cuda_stream)
        done = self.s_comp.record_event()
        self.s_prep.wait_event(done)  # the region's memory is released on its own (prepare) stream
        with torch.cuda.stream(self.s_prep):
            region.close()
        self.regions += 1

    def finish(self):
        """Drain the pipeline; afterwards the caller's current stream sees every accumulator update."""
        if self._pending is not None:
            self._compute(self._pending)
            self._pending = None
        self.main.wait_stream(self.s_comp)
        self.main.wait_stream(self.s_prep)
        self.s_comp.synchronize()
        self.s_prep.synchronize()
        self._reap(everything=True)

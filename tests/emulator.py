"""numpy emulation of the C ABI (`include/pileup_b200.h`) -- TEST INFRASTRUCTURE ONLY.

It mirrors what the CUDA kernels compute, *including* the raw accumulator layout and the `num` decomposition
(`n_fast - rowbad - colbad + tile`), so that the host pipeline
(`coolpuppy_b200/coolpup.py`: window generation, slot assignment, flip merging, final normalisation, DataFrame
schema) and the decomposition itself can be checked against the golden vectors on a machine without a GPU.
Tests install it with `monkeypatch` over `coolpuppy_b200._native`; the product never imports this module.
"""
import numpy as np

from coolpuppy_b200 import _native

F_OOE, F_EXPCTRL, F_COVERAGE, F_NODIAG, F_LOCAL = 1, 2, 4, 8, 32


def zoom_array(in_array, final_shape):
    """cooltools.lib.numutils.zoom_array with the real scipy.ndimage.zoom(order=1) (see oracle/refshim)."""
    from scipy.ndimage import zoom

    in_array = np.asarray(in_array, dtype=np.double)
    mults = [int(np.ceil(i / f)) if f < i else 1 for i, f in zip(in_array.shape, final_shape)]
    temp_shape = tuple(f * m for f, m in zip(final_shape, mults))
    rescaled = zoom(in_array, np.array(temp_shape) / np.array(in_array.shape) + 0.0000001, order=1)
    for ind, mult in enumerate(mults):
        if mult != 1:
            sh = list(rescaled.shape)
            rescaled.shape = sh[:ind] + [sh[ind] // mult, mult] + sh[ind + 1 :]
            rescaled = np.mean(rescaled, axis=ind + 1)
    return rescaled


def layout(W):
    w2 = W * W
    L = dict(w2=w2, num=w2, rb=2 * w2)
    L["cb"] = L["rb"] + W
    L["covs"] = L["cb"] + W
    L["cove"] = L["covs"] + W
    L["tsum"] = L["cove"] + W
    L["tnum"] = L["tsum"] + 2 * W
    L["n"] = L["tnum"] + 2 * W
    L["nfast"] = L["n"] + 1
    L["stride"] = L["n"] + 8
    return L


def symmetrise(nb, indptr_u, col_u, cnt_u):
    row = np.repeat(np.arange(nb), np.diff(indptr_u))
    # the device code requires sorted columns within every row (it mirrors by a stable sort and takes the
    # in-region columns as a prefix): enforce the same contract here
    if col_u.size > 1:
        same_row = row[1:] == row[:-1]
        assert np.all(np.diff(col_u.astype(np.int64))[same_row] > 0), "upper CSR rows must have sorted, unique columns"
    keep = (col_u >= 0) & (col_u < nb)
    row, col, cnt = row[keep], col_u[keep].astype(np.int64), cnt_u[keep]
    off = row != col
    r = np.concatenate([row, col[off]])
    c = np.concatenate([col, row[off]])
    v = np.concatenate([cnt, cnt[off]])
    order = np.lexsort((c, r))
    ip = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=nb), out=ip[1:])
    return ip.astype(np.int32), c[order].astype(np.int32), v[order].astype(np.int32)


class EmuRegion:
    def __init__(self, device, nb, indptr, col, count, weight=None, expected=None, coverage=None, ignore_diags=2,
                 flags=0, stream=0, upper=False):
        if upper:  # mirror the stored upper triangle like pup_region_create_upper
            indptr, col, count = symmetrise(int(nb), np.asarray(indptr), np.asarray(col), np.asarray(count))
        self.nb = int(nb)
        self.ignore_diags = int(ignore_diags)
        self.region_flags = int(flags) & (F_OOE | F_NODIAG)
        self.indptr, self.col, self.count = np.asarray(indptr), np.asarray(col), np.asarray(count)
        self.weight, self.expected, self.coverage = weight, expected, coverage
        self.bad = np.isnan(weight) if weight is not None else np.zeros(nb, dtype=bool)
        if expected is not None:
            self.ebad = np.isnan(expected) | (expected == 0)
        else:
            self.ebad = np.zeros(nb, dtype=bool)
        self.ebadpre = np.concatenate([[0], np.cumsum(self.ebad)])
        self.balanced = weight is not None
        self.device_bytes = 0

    def _slow(self, r0, c0, W, igd, flags):
        D0 = c0 - r0
        dmin, dmax = D0 - (W - 1), D0 + (W - 1)
        if not (flags & F_NODIAG) and dmin < igd:
            return True
        if flags & F_OOE:
            if dmin >= 0:
                a, b = dmin, dmax
            elif dmax <= 0:
                a, b = -dmax, -dmin
            else:
                a, b = 0, max(-dmin, dmax)
            b = min(b, self.nb - 1)
            a = min(a, b)
            return self.ebadpre[b + 1] - self.ebadpre[a] > 0
        return False

    def accumulate(self, r0, c0, slot, W, n_slots, flags, acc, stream=0, want_n_valid=False):
        ignore_diags = self.ignore_diags
        flags = (int(flags) & (F_EXPCTRL | F_COVERAGE)) | self.region_flags
        L = layout(W)
        a = acc.numpy() if hasattr(acc, "numpy") else acc
        a = a.reshape(n_slots, L["stride"])
        nv = 0
        nb = self.nb
        m = np.arange(-(W - 1), W)
        for i in range(len(r0)):
            r, c, s = int(r0[i]), int(c0[i]), int(slot[i])
            if r < 0 or c < 0 or r + W > nb or c + W > nb or s < 0 or s >= n_slots:
                continue
            nv += 1
            A = a[s]
            slow = self._slow(r, c, W, ignore_diags, flags)
            A[L["n"]] += 1
            if not slow:
                A[L["nfast"]] += 1
                A[L["rb"] : L["rb"] + W] += self.bad[r : r + W]
                A[L["cb"] : L["cb"] + W] += self.bad[c : c + W]
            if flags & F_COVERAGE:
                A[L["covs"] : L["covs"] + W] += np.nan_to_num(self.coverage[r : r + W], nan=0.0, posinf=np.inf, neginf=-np.inf)
                A[L["cove"] : L["cove"] + W] += np.nan_to_num(self.coverage[c : c + W], nan=0.0, posinf=np.inf, neginf=-np.inf)
            if flags & F_EXPCTRL:
                e = self.expected[np.abs(c - r + m)]
                A[L["tsum"] : L["tsum"] + 2 * W - 1] += np.where(np.isnan(e), 0.0, e)
                A[L["tnum"] : L["tnum"] + 2 * W - 1] += np.isfinite(e)
            S = A[: L["w2"]].reshape(W, W)
            N = A[L["num"] : L["num"] + L["w2"]].reshape(W, W)
            for di in range(W):
                row = r + di
                rbad = bool(self.bad[row])
                if not rbad:
                    lo, hi = self.indptr[row], self.indptr[row + 1]
                    cols = self.col[lo:hi]
                    k0, k1 = np.searchsorted(cols, c), np.searchsorted(cols, c + W)
                    cc = cols[k0:k1]
                    v = self.count[lo + k0 : lo + k1].astype(np.float64)
                    d = cc - row
                    keep = np.ones(len(cc), dtype=bool) if (flags & F_NODIAG) else d >= ignore_diags
                    if self.balanced:
                        v = (self.weight[row] * self.weight[cc]) * v
                    if flags & F_OOE:
                        with np.errstate(divide="ignore", invalid="ignore"):
                            v = v / self.expected[np.abs(d)]
                    keep &= ~np.isnan(v)
                    np.add.at(S[di], cc[keep] - c, v[keep])
                if slow:
                    cs = np.arange(c, c + W)
                    d = cs - row
                    ok = np.full(W, not rbad) & ~self.bad[cs]
                    if not (flags & F_NODIAG):
                        ok &= d >= ignore_diags
                    if flags & F_OOE:
                        ok &= ~self.ebad[np.abs(d)]
                    N[di] += ok
                elif rbad:
                    N[di] += self.bad[c : c + W]
        return nv if want_n_valid else None

    def _snippet(self, r, c, h, w):
        """Dense snippet with the reference's NaN semantics (coolpup.py:1115-1156)."""
        from scipy import sparse

        if not hasattr(self, "_mat"):
            self._mat = sparse.csr_matrix((self.count.astype(np.float64), self.col, self.indptr), shape=(self.nb, self.nb))
        data = self._mat[r : r + h, c : c + w].toarray().astype(float)
        if self.balanced:
            data = (self.weight[r : r + h, None] * self.weight[None, c : c + w]) * data
        data[self.bad[r : r + h], :] = np.nan
        data[:, self.bad[c : c + w]] = np.nan
        ii, jj = np.arange(r, r + h)[:, None], np.arange(c, c + w)[None, :]
        if not (self.region_flags & F_NODIAG):
            data[(jj - ii) < self.ignore_diags] = np.nan
        if self.region_flags & F_OOE:
            with np.errstate(divide="ignore", invalid="ignore"):
                data = data / self.expected[np.abs(jj - ii)]
        return data

    def accumulate_rescaled(self, r0, c0, h, w, slot, mode, rs, n_slots, flags, acc, stream=0, want_n_valid=False):
        """_rescale_snip (coolpup.py:1193-1234) per window + accumulate, with scipy's zoom."""
        import warnings

        L = layout(rs)
        a = (acc.numpy() if hasattr(acc, "numpy") else acc).reshape(n_slots, L["stride"])
        nv = 0
        for i in range(len(r0)):
            r, c, hh, ww, s = int(r0[i]), int(c0[i]), int(h[i]), int(w[i]), int(slot[i])
            if r < 0 or c < 0 or r + hh > self.nb or c + ww > self.nb or s < 0 or s >= n_slots:
                continue
            nv += 1
            if mode is not None and int(mode[i]) == 1:
                ii, jj = np.arange(r, r + hh)[:, None], np.arange(c, c + ww)[None, :]
                data = self.expected[np.abs(jj - ii)].astype(float)
            else:
                data = self._snippet(r, c, hh, ww)
            if data.size == 0 or np.all(np.isnan(data)):
                data = np.zeros((rs, rs))
            else:
                if int(flags) & F_LOCAL:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore", category=RuntimeWarning)
                        data = np.nanmean(np.dstack((data, data.T)), 2)
                nans = np.isnan(data) * 1
                data = zoom_array(np.nan_to_num(data), (rs, rs))
                nanzoom = zoom_array(nans, (rs, rs))
                data[np.ceil(nanzoom).astype(bool)] = np.nan
            A = a[s]
            A[L["n"]] += 1
            S = A[: L["w2"]].reshape(rs, rs)
            N = A[L["num"] : L["num"] + L["w2"]].reshape(rs, rs)
            fin = np.isfinite(data)
            S += np.where(np.isnan(data), 0.0, data)
            N += fin
            if int(flags) & F_COVERAGE and hh > 0 and ww > 0:
                A[L["covs"] : L["covs"] + rs] += np.nan_to_num(zoom_array(self.coverage[r : r + hh], (rs,)), nan=0.0)
                A[L["cove"] : L["cove"] + rs] += np.nan_to_num(zoom_array(self.coverage[c : c + ww], (rs,)), nan=0.0)
        return nv if want_n_valid else None

    def stripes(self, r0, c0, W, stream=0):
        n = len(r0)
        hor = np.full((n, W), np.nan)
        ver = np.full((n, W), np.nan)
        cn = W // 2
        flags, igd = self.region_flags, self.ignore_diags

        def pixel(r, c):
            if self.bad[r] or self.bad[c]:
                return np.nan
            d = c - r
            if not (flags & F_NODIAG) and d < igd:
                return np.nan
            lo, hi = self.indptr[r], self.indptr[r + 1]
            k = lo + np.searchsorted(self.col[lo:hi], c)
            stored = k < hi and self.col[k] == c
            v = float(self.count[k]) if stored else 0.0
            if self.balanced:
                v = (self.weight[r] * self.weight[c]) * v
            if flags & F_OOE:
                with np.errstate(divide="ignore", invalid="ignore"):
                    v = np.float64(v) / self.expected[abs(d)]
            return v

        for i in range(n):
            r, c = int(r0[i]), int(c0[i])
            if r < 0 or c < 0 or r + W > self.nb or c + W > self.nb:
                continue
            for k in range(W):
                hor[i, k] = pixel(r + cn, c + k)
                ver[i, k] = pixel(r + (W - 1 - k), c + cn)
        return hor, ver

    def algorithmic_bytes(self, r0, c0, W, flags=0, stream=0):
        return 0, 0

    def close(self):
        pass


def emu_export(acc, W, n_slots, device=0, stream=0, want_expected=False, want_cov=False):
    L = layout(W)
    a = (acc.numpy() if hasattr(acc, "numpy") else acc).reshape(n_slots, L["stride"])
    if n_slots == 0:
        return {"sum": np.zeros((0, W, W)), "num": np.zeros((0, W, W), dtype=np.int64), "n": np.zeros(0, dtype=np.int64)}
    out = {"sum": a[:, : L["w2"]].reshape(n_slots, W, W).copy(), "n": np.rint(a[:, L["n"]]).astype(np.int64)}
    numt = a[:, L["num"] : L["num"] + L["w2"]].reshape(n_slots, W, W)
    rb = a[:, L["rb"] : L["rb"] + W]
    cb = a[:, L["cb"] : L["cb"] + W]
    out["num"] = np.rint(a[:, L["nfast"]][:, None, None] - rb[:, :, None] - cb[:, None, :] + numt).astype(np.int64)
    if want_cov:
        out["cov_start"] = a[:, L["covs"] : L["covs"] + W].copy()
        out["cov_end"] = a[:, L["cove"] : L["cove"] + W].copy()
    if want_expected:
        i = np.arange(W)
        idx = i[None, :] - i[:, None] + W - 1
        out["exp_sum"] = a[:, L["tsum"] : L["tsum"] + 2 * W][:, idx]
        out["exp_num"] = np.rint(a[:, L["tnum"] : L["tnum"] + 2 * W][:, idx]).astype(np.int64)
    return out


class SerialPipeline:
    """Stand-in for coolpuppy_b200.pipeline.RegionPipeline: same interface, regions done one after the other."""

    def __init__(self, device, W, n_slots, flags):
        self.W, self.n_slots, self.flags = W, n_slots, flags
        self.launches = 0
        self.regions = 0

    def submit(self, region_kwargs, windows, acc, after=None, windows_on_device=None):
        region = EmuRegion(0, **region_kwargs)
        r0, c0, slot = windows[:3]
        if len(windows) >= 5:
            mode = np.asarray(windows[5]) if len(windows) == 6 else None
            region.accumulate_rescaled(np.asarray(r0), np.asarray(c0), np.asarray(windows[3]), np.asarray(windows[4]),
                                       np.asarray(slot), mode, self.W, self.n_slots, self.flags, acc)
        else:
            region.accumulate(np.asarray(r0), np.asarray(c0), np.asarray(slot), self.W, self.n_slots, self.flags, acc)
        if after is not None:
            after(region, 0)
        self.regions += 1
        return None

    def finish(self):
        pass


def install(monkeypatch):
    """Route coolpuppy_b200._native's device entry points to the emulator (tests only)."""
    import torch

    monkeypatch.setattr(_native, "Region", EmuRegion)
    monkeypatch.setattr(_native, "acc_export", emu_export)
    monkeypatch.setattr(_native, "require_device", lambda: None)
    monkeypatch.setattr(_native, "alloc_accumulator", lambda n, device: torch.zeros(int(n), dtype=torch.float64))
    monkeypatch.setattr(_native, "current_stream", lambda device: 0)
    monkeypatch.setattr(_native, "acc_stride", lambda W: layout(int(W))["stride"])
    monkeypatch.setattr(_native, "make_pipeline", SerialPipeline)
    monkeypatch.setattr(_native, "device_windows_supported", lambda: False)

    real = _native.lib()  # host-side entry points (window layout) are the real library: they need no device

    class _L:
        @staticmethod
        def pup_last_launches():
            return 0

        def __getattr__(self, name):
            if name.startswith("pup_pair_windows") or name == "pup_last_error":
                return getattr(real, name)
            raise AttributeError(name)

    stub = _L()
    monkeypatch.setattr(_native, "lib", lambda: stub)

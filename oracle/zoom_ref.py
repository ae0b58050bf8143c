"""TEST INFRASTRUCTURE ONLY (see oracle/pileup_oracle.py): a loop-level restatement of the arithmetic of
``scipy.ndimage.zoom(a, z, order=1)`` (mode="constant", cval=0, grid_mode=False -- the call made by
``cooltools.lib.numutils.zoom_array``, which ``PileUpper._rescale_snip`` uses, coolpup.py:1223-1233).

``k_rescale`` (coolpuppy_b200/csrc/pileup_b200.cu) mirrors exactly these operations, in this order, with
round-to-nearest fp64 and no fused multiply-add.  scipy itself is a third-party dependency of the reference (unpinned in
its requirements.txt); the installed scipy is the pin: ``tests/test_zoom_ref.py`` compares this restatement with
``scipy.ndimage.zoom`` bit for bit, and the emulator (tests/emulator.py) and the goldens use the real scipy.

Algorithm (scipy/ndimage/src/ni_interpolation.c::NI_ZoomShift with order 1, as published):
  * output length ``m`` along an axis of input length ``n``; scale ``zf = (n - 1) / (m - 1)`` (1 when m == 1);
  * output sample k reads coordinate ``cc = k * zf``; ``cc > n - 1`` (rounding) or ``cc < 0`` -> the constant 0;
  * ``i0 = floor(cc)``, ``x = cc - i0``, weights ``w0 = 1 - x``, ``w1 = 1 - w0`` (the last spline weight is one minus
    the sum of the others), ``i1 = i0 + 1`` reflected at the edge (``2n - 2 - i1``; its weight is 0 there);
  * 2-D: ``sum over (a, b) in ((0,0), (0,1), (1,0), (1,1)) of (D[ia, jb] * wrow_a) * wcol_b``, added in that order.
"""
import numpy as np


def zoom_plan(n, m):
    """(i0, i1, w0, w1) per output sample; w0 = w1 = 0 marks the constant-0 sample."""
    i0 = np.zeros(m, dtype=np.int64)
    i1 = np.zeros(m, dtype=np.int64)
    w0 = np.zeros(m)
    w1 = np.zeros(m)
    zf = (n - 1) / (m - 1) if m > 1 else 1.0
    for k in range(m):
        cc = k * zf
        if cc < 0.0 or cc > n - 1:
            continue
        fl = np.floor(cc)
        x = cc - fl
        w0[k] = 1.0 - x
        w1[k] = 1.0 - w0[k]
        if n > 1:
            i0[k] = int(fl)
            i1[k] = i0[k] + 1
            if i1[k] >= n:
                i1[k] = 2 * n - 2 - i1[k]
    return i0, i1, w0, w1


def zoom_linear_2d(D, out_shape):
    """scipy.ndimage.zoom(D, ., order=1) to ``out_shape`` (2-D)."""
    D = np.asarray(D, dtype=np.float64)
    ri0, ri1, rw0, rw1 = zoom_plan(D.shape[0], out_shape[0])
    ci0, ci1, cw0, cw1 = zoom_plan(D.shape[1], out_shape[1])
    out = np.zeros(out_shape)
    for a in range(out_shape[0]):
        if rw0[a] == 0.0 and rw1[a] == 0.0:
            continue
        for b in range(out_shape[1]):
            if cw0[b] == 0.0 and cw1[b] == 0.0:
                continue
            t = (D[ri0[a], ci0[b]] * rw0[a]) * cw0[b]
            t = t + (D[ri0[a], ci1[b]] * rw0[a]) * cw1[b]
            t = t + (D[ri1[a], ci0[b]] * rw1[a]) * cw0[b]
            t = t + (D[ri1[a], ci1[b]] * rw1[a]) * cw1[b]
            out[a, b] = t
    return out


def zoom_linear_1d(v, m):
    v = np.asarray(v, dtype=np.float64)
    i0, i1, w0, w1 = zoom_plan(v.shape[0], m)
    out = np.zeros(m)
    for k in range(m):
        if w0[k] != 0.0 or w1[k] != 0.0:
            out[k] = v[i0[k]] * w0[k] + v[i1[k]] * w1[k]
    return out


def zoom_array_ref(D, rs):
    """``zoom_array(D, (rs, rs))`` the way ``k_rescale`` evaluates it: interpolate to ``rs * mult`` per axis, then block
    means -- rows first, then columns, each a sequential sum divided by the block length (``np.mean`` sums the row
    blocks sequentially; along the contiguous column axis it switches to pairwise summation for blocks of 8 or more,
    which can differ from the sequential sum in the last bit)."""
    D = np.asarray(D, dtype=np.float64)
    mr = int(np.ceil(D.shape[0] / rs)) if D.shape[0] > rs else 1
    mc = int(np.ceil(D.shape[1] / rs)) if D.shape[1] > rs else 1
    Z = zoom_linear_2d(D, (rs * mr, rs * mc))
    out = np.zeros((rs, rs))
    for i in range(rs):
        for j in range(rs):
            acc = 0.0
            for bj in range(mc):
                col = 0.0
                for bi in range(mr):
                    col = col + Z[i * mr + bi, j * mc + bj]
                if mr > 1:
                    col = col / mr
                acc = acc + col
            if mc > 1:
                acc = acc / mc
            out[i, j] = acc
    return out

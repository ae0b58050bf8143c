"""CPU tests of the C ABI surface: the library loads, exports everything include/pileup_b200.h declares, and the
compute path fails loudly without a GPU (there is no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from coolpuppy_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pileup_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pup_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _native.lib()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_native.SYMBOLS)
    assert lib.pup_abi_version() >= 2


def test_acc_stride_and_flags():
    for W in (5, 21, 83, 203):
        assert _native.acc_stride(W) == 2 * W * W + 8 * W + 8
    assert (_native.PUP_F_OOE, _native.PUP_F_EXPCTRL, _native.PUP_F_COVERAGE, _native.PUP_F_NODIAG, _native.PUP_F_ASYNC) == (1, 2, 4, 8, 16)


def test_no_cpu_fallback_without_gpu():
    if _native.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_native.NativeError):
        _native.require_device()
    ip = np.zeros(11, dtype=np.int32)
    z = np.zeros(0, dtype=np.int32)
    with pytest.raises(_native.NativeError):
        _native.Region(0, 10, ip, z, z)
    acc = np.zeros(_native.acc_stride(5))
    with pytest.raises(_native.NativeError):
        _native.accumulate_region(0, 10, ip, z, z, None, None, None, z, z, z, 5, 2, 1, 0, acc)
    import golden_util as gu
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs("toy_strand_balanced")
    with pytest.raises(_native.NativeError):
        cp.pileup(clr, feats, **kw)


def test_bad_arguments_are_rejected_before_any_device_work():
    lib = _native.lib()
    assert lib.pup_acc_export(None, 5, 1, 0, None, None, None, None, None, None, None, None) == -1
    assert b"bad arguments" in lib.pup_last_error()
    assert lib.pup_accumulate(None, 0, None, None, None, 5, 1, 0, None, None, None) == -1
    # host-side decode of an accumulator buffer needs no device
    W, n_slots = 3, 2
    L = _native.acc_stride(W)
    acc = np.zeros(n_slots * L)
    acc[0:9] = np.arange(9)            # sum of slot 0
    acc[2 * 9 + 8 * W] = 4             # n
    acc[2 * 9 + 8 * W + 1] = 3         # n_fast
    acc[2 * 9 + 0] = 1                 # rb[0]
    acc[2 * 9 + W + 2] = 2             # cb[2]
    acc[9 + 2] = 1                     # xtile[0][2]
    out = _native.acc_export(acc, W, n_slots)
    assert out["n"].tolist() == [4, 0]
    assert np.array_equal(out["sum"][0], np.arange(9.0).reshape(3, 3))
    assert out["num"][0].tolist() == [[2, 2, 1], [3, 3, 1], [3, 3, 1]]


def test_pair_window_layout_host_code_matches_numpy():
    """pup_pair_windows_count / _fill (plain host code, no device) against a direct numpy restatement of the pair loop
    of CoordCreator.get_combinations + _control_regions (coolpup.py:682-714, 387-453)."""
    rng = np.random.default_rng(5)
    for m, nctrl in [(0, 2), (1, 0), (2, 3), (37, 0), (60, 4)]:
        center = np.sort(rng.integers(0, 5_000_000, m)).astype(np.float64) + 0.5
        stbin = (center // 10_000).astype(np.int64) - 10
        mind, maxd = 300_000.0, 2_500_000.0
        q, total = _native.pair_windows_count(center, mind, maxd)
        exp_k, exp_l = [], []
        for i in range(1, m):
            k = np.arange(m - i)
            d = np.abs(center[k + i] - center[k])
            kk = k[(mind <= d) & (d <= maxd)]
            assert q[i] == len(kk)
            exp_k.append(kk)
            exp_l.append(kk + i)
        assert total == sum(len(x) for x in exp_k)
        dbin = rng.integers(-100, 100, total * nctrl).astype(np.int64) if nctrl else None
        st1, st2, kind, i1, i2, dist = _native.pair_windows_fill(stbin, center, mind, maxd, nctrl, dbin, total)
        e1, e2, ek, ekk, ell = [], [], [], [], []
        pos = 0
        for kk, ll in zip(exp_k, exp_l):
            n = len(kk)
            sh = np.concatenate([np.zeros(n, dtype=np.int64), dbin[pos : pos + n * nctrl]]) if nctrl else np.zeros(n, dtype=np.int64)
            pos += n * nctrl
            rk, rl = np.tile(kk, nctrl + 1), np.tile(ll, nctrl + 1)
            e1.append(stbin[rk] + sh)
            e2.append(stbin[rl] + sh)
            kd = np.ones(n * (nctrl + 1), dtype=np.int8)
            kd[:n] = 0
            ek.append(kd)
            ekk.append(rk)
            ell.append(rl)
        cat = lambda x, dt: np.concatenate(x).astype(dt) if x else np.zeros(0, dtype=dt)
        assert np.array_equal(st1, cat(e1, np.int64)) and np.array_equal(st2, cat(e2, np.int64))
        assert np.array_equal(kind, cat(ek, np.int8))
        assert np.array_equal(i1, cat(ekk, np.int64)) and np.array_equal(i2, cat(ell, np.int64))
        assert np.array_equal(dist, center[i2] - center[i1])

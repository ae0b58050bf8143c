"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI."""
import warnings

import numpy as np
import pytest

import golden_util as gu
from synth import random_region, random_windows

pytestmark = pytest.mark.gpu

RTOL = 1e-6  # BASELINE.json north_star: "within 1e-6 relative"; observed differences are ~1e-15 (fp64 sum order)


def _cuda():
    from coolpuppy_b200 import _native

    if _native.device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return _native


def _compare_rows(pups, name):
    from oracle.pileup_oracle import key_repr

    z, _ = gu.load_golden(name)
    if "group" in pups.columns:
        keys = [key_repr(g) for g in pups["group"]]
    else:
        keys = [repr((r.chrom, int(r.start), int(r.end))) for r in pups.itertuples()]
    assert keys == [str(k) for k in z["row_keys"]], "row keys / order differ from the reference"
    assert list(pups.columns) == [str(c) for c in z["columns"]], "DataFrame schema differs from the reference"
    for i in range(len(keys)):
        g = {f.split(".", 1)[1]: z[f] for f in z.files if f.startswith(f"row{i}.")}
        a = np.asarray(pups["data"].iloc[i], dtype=float)
        b = g["data"]
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{name} row {keys[i]}: NaN pattern differs"
        m = np.isfinite(b)
        np.testing.assert_allclose(a[m], b[m], rtol=RTOL, atol=0)
        assert int(pups["n"].iloc[i]) == int(g["n"])
        assert np.array_equal(np.asarray(pups["num"].iloc[i]), g["num"])
        if "control_n" in g:
            assert int(pups["control_n"].iloc[i]) == int(g["control_n"])
            assert np.array_equal(np.asarray(pups["control_num"].iloc[i]), g["control_num"])
        if "vertical_stripe" in g:
            for f in ("vertical_stripe", "horizontal_stripe"):
                np.testing.assert_allclose(np.asarray(pups[f].iloc[i], dtype=float), g[f], rtol=RTOL, equal_nan=True)
            assert np.array_equal(np.asarray(pups["coordinates"].iloc[i]).astype(str), g["coordinates"])


@pytest.mark.parametrize("name", gu.all_cases())
def test_golden_case_through_cuda(name):
    """coolpuppy_b200.pileup() on the GPU == the real reference's stored output (tests/golden)."""
    _cuda()
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pups = cp.pileup(clr, feats, **kw)
    _compare_rows(pups, name)


MODES = {
    "raw": dict(bal=False, flags=0),
    "balanced": dict(bal=True, flags=0),
    "balanced_ooe": dict(bal=True, flags=1, exp=True),
    "raw_ooe": dict(bal=False, flags=1, exp=True),
    "balanced_expctrl": dict(bal=True, flags=2, exp=True),
    "raw_coverage": dict(bal=False, flags=4, cov=True),
}


def _check_against_oracle(nv, out, ref):
    assert nv == int(ref["n"].sum())
    assert np.array_equal(out["n"], ref["n"])
    assert np.array_equal(out["num"], ref["num"])
    assert np.array_equal(np.isinf(out["sum"]), np.isinf(ref["sum"]))
    m = np.isfinite(ref["sum"])
    np.testing.assert_allclose(out["sum"][m], ref["sum"][m], rtol=RTOL, atol=1e-300)


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("nb,W,dens,nwin,n_slots", [(300, 21, 30, 500, 3), (700, 83, 200, 300, 2), (64, 5, 4, 200, 7),
                                                      (900, 203, 400, 60, 2), (2000, 21, 3, 3000, 40)])
@pytest.mark.parametrize("memory", ["host", "device"])
def test_kernel_vs_oracle_random(mode, nb, W, dens, nwin, n_slots, memory):
    """pup_accumulate on seeded random CSR + windows == the dense per-window restatement, host and device buffers."""
    nat = _cuda()
    from oracle.pileup_oracle import oracle_accumulate

    cfg = MODES[mode]
    ip, col, cnt, w, e, cov = random_region(nb, dens, seed=nb + W, nan_frac=0.05, with_expected=True, with_cov=True)
    weight = w if cfg["bal"] else None
    expected = e if cfg.get("exp") else None
    coverage = cov if cfg.get("cov") else None
    r0, c0, sl = random_windows(nb, W, nwin, n_slots, seed=7 * nb + W)
    ref = oracle_accumulate(nb, ip, col, cnt, weight, expected, coverage, r0, c0, sl, W, 2, n_slots,
                            ooe=cfg["flags"] == 1, expctrl=cfg["flags"] == 2)
    stride = nat.acc_stride(W)
    if memory == "host":
        acc = np.zeros(n_slots * stride)
        nv = nat.accumulate_region(0, nb, ip, col, cnt, weight, expected, coverage, r0, c0, sl, W, 2, n_slots,
                                   cfg["flags"], acc)
        out = nat.acc_export(acc, W, n_slots, want_expected=True, want_cov=True)
    else:
        import torch

        dev = torch.device("cuda", 0)
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        tens = [t(x) for x in (ip, col, cnt, weight, expected, coverage, r0, c0, sl)]
        acc = torch.zeros(n_slots * stride, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        reg = nat.Region(0, nb, tens[0], tens[1], tens[2], tens[3], tens[4], tens[5], ignore_diags=2,
                         flags=cfg["flags"], stream=stream)
        # two calls into the same accumulator: the buffer is purely additive
        half = nwin // 2
        nv = reg.accumulate(tens[6][:half].contiguous(), tens[7][:half].contiguous(), tens[8][:half].contiguous(), W,
                            n_slots, cfg["flags"], acc, stream=stream, want_n_valid=True)
        nv += reg.accumulate(tens[6][half:].contiguous(), tens[7][half:].contiguous(), tens[8][half:].contiguous(), W,
                             n_slots, cfg["flags"], acc, stream=stream, want_n_valid=True)
        out = nat.acc_export(acc, W, n_slots, device=0, stream=stream, want_expected=True, want_cov=True)
        reg.close()
    _check_against_oracle(nv, out, ref)
    if cfg["flags"] == 2:
        assert np.array_equal(out["exp_num"], ref["exp_num"])
        np.testing.assert_allclose(out["exp_sum"], ref["exp_sum"], rtol=RTOL)
    if cfg.get("cov"):
        np.testing.assert_allclose(out["cov_start"], ref["cov_start"], rtol=RTOL)
        np.testing.assert_allclose(out["cov_end"], ref["cov_end"], rtol=RTOL)


@pytest.mark.parametrize("strip,lanes", [(1, 4), (1, 8), (2, 8), (2, 16), (4, 8), (4, 16), (8, 16), (8, 32)])
@pytest.mark.parametrize("nb,W,dens,nwin,n_slots", [(701, 83, 200, 400, 3), (333, 21, 40, 700, 5), (900, 203, 400, 50, 2)])
def test_strip_geometries(monkeypatch, strip, lanes, nb, W, dens, nwin, n_slots):
    """Every strip height R / lane count S the main kernel is instantiated for gives the oracle's accumulators
    (the default is R = 2, S = 8; the geometry is fixed when the region is created)."""
    nat = _cuda()
    from oracle.pileup_oracle import oracle_accumulate

    monkeypatch.setenv("PUP_STRIP", str(strip))
    monkeypatch.setenv("PUP_LANES", str(lanes))
    ip, col, cnt, w, e, cov = random_region(nb, dens, seed=nb + W, nan_frac=0.05, with_expected=True)
    r0, c0, sl = random_windows(nb, W, nwin, n_slots, seed=3 * nb + W)
    ref = oracle_accumulate(nb, ip, col, cnt, w, e, None, r0, c0, sl, W, 2, n_slots, ooe=True)
    acc = np.zeros(n_slots * nat.acc_stride(W))
    nv = nat.accumulate_region(0, nb, ip, col, cnt, w, e, None, r0, c0, sl, W, 2, n_slots, nat.PUP_F_OOE, acc)
    _check_against_oracle(nv, nat.acc_export(acc, W, n_slots), ref)
    # the per-window stripes read the same strip layout
    reg = nat.Region(0, nb, ip, col, cnt, w, e, None, ignore_diags=2, flags=nat.PUP_F_OOE)
    hor, ver = reg.stripes(r0[:40], c0[:40], W)
    b, z = reg.algorithmic_bytes(r0, c0, W)
    reg.close()
    from scipy import sparse

    from emulator import EmuRegion

    mat = sparse.csr_matrix((cnt, col, ip), shape=(nb, nb))
    inb = (r0 >= 0) & (c0 >= 0) & (r0 + W <= nb) & (c0 + W <= nb)
    # algorithmic pixels = stored pixels inside the window that the signed diagonal mask keeps (col - row >= 2);
    # masked pixels are not stored on the device at all
    def kept(a, c):
        blk = mat[a : a + W, c : c + W].tocoo()
        return int(((blk.col + c) - (blk.row + a) >= 2).sum())

    assert z == sum(kept(a, c) for a, c in zip(r0[inb], c0[inb]))
    ehor, ever = EmuRegion(0, nb, ip, col, cnt, w, e, None, ignore_diags=2, flags=nat.PUP_F_OOE).stripes(r0[:40], c0[:40], W)
    np.testing.assert_allclose(hor, ehor, rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(ver, ever, rtol=RTOL, equal_nan=True)


@pytest.mark.parametrize("env", [{"PUP_SCHED": "0"}, {"PUP_SCHED": "1"}, {"PUP_TILE_PAD": "1"},
                                 {"PUP_TILE_PAD": "5", "PUP_TILE_INTERLEAVE": "0"}, {"PUP_TILE_INTERLEAVE": "1"},
                                 {"PUP_TILE_INTERLEAVE": "0", "PUP_CHUNK": "16"}, {"PUP_SORT_C0_BITS": "0"}, {"PUP_SORT_C0_BITS": "3"}, {"PUP_CHUNK": "7"},
                                 {"PUP_DENSE": "0"}, {"PUP_BAND": "0"}, {"PUP_BAND_PCT": "3000", "PUP_BAND_DENSITY_PCT": "1"},
                                 {"PUP_BAND_PCT": "50", "PUP_BAND_DENSITY_PCT": "60"}])
@pytest.mark.parametrize("nb,W,dens,nwin,n_slots", [(700, 83, 200, 2500, 3), (900, 203, 400, 90, 2), (300, 21, 30, 4000, 5)])
def test_main_kernel_variants(monkeypatch, env, nb, W, dens, nwin, n_slots):
    """Scheduling (static round-robin / barrier-free dynamic ring), occupancy and tile-layout variants of the main
    kernel all give the oracle's accumulators."""
    nat = _cuda()
    from oracle.pileup_oracle import oracle_accumulate

    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ip, col, cnt, w, e, cov = random_region(nb, dens, seed=nb + W + 1, nan_frac=0.05, with_expected=True)
    r0, c0, sl = random_windows(nb, W, nwin, n_slots, seed=5 * nb + W)
    ref = oracle_accumulate(nb, ip, col, cnt, w, e, None, r0, c0, sl, W, 2, n_slots, ooe=True)
    acc = np.zeros(n_slots * nat.acc_stride(W))
    nv = nat.accumulate_region(0, nb, ip, col, cnt, w, e, None, r0, c0, sl, W, 2, n_slots, nat.PUP_F_OOE, acc)
    _check_against_oracle(nv, nat.acc_export(acc, W, n_slots), ref)


@pytest.mark.parametrize("strip,lanes", [(2, 8), (4, 8), (1, 4), (8, 32)])
def test_adversarial_duplicate_windows(monkeypatch, strip, lanes):
    """The tile read-modify-write has no atomics: correctness rests on (a) one lane group owning its tile rows and
    (b) the windows a group has in flight being added one after the other.  Worst cases for both: thousands of
    IDENTICAL windows in one slot (every window in flight hits the same cells in the same half-step), windows that
    differ only in c0 by 1 (neighbouring lanes / windows hit neighbouring and identical cells), and a dense block
    so that every lane of a group holds a pixel in every half-step.  Integer counts and power-of-two weights make
    the expected sums exact, so any lost update shows up as an inequality, not a tolerance failure."""
    nat = _cuda()
    monkeypatch.setenv("PUP_STRIP", str(strip))
    monkeypatch.setenv("PUP_LANES", str(lanes))
    from scipy import sparse

    nb, W = 400, 21
    rng = np.random.default_rng(42)
    dense = sparse.random(nb, nb, density=0.9, random_state=7, data_rvs=lambda n: rng.integers(1, 9, n)).tocsr()
    dense = sparse.triu(dense) + sparse.triu(dense, 1).T
    dense = sparse.csr_matrix(dense)
    dense.sort_indices()
    ip, col, cnt = dense.indptr.astype(np.int32), dense.indices.astype(np.int32), dense.data.astype(np.int32)
    r0 = np.concatenate([np.full(4096, 100), np.full(4, 37), np.full(1000, 101), np.arange(60, 64)]).astype(np.int32)
    c0 = np.concatenate([np.full(4096, 230), 300 + np.arange(4), np.full(1000, 230), np.full(4, 200)]).astype(np.int32)
    sl = np.concatenate([np.zeros(4096), np.ones(4), np.zeros(1000), np.ones(4)]).astype(np.int32)
    full = dense.toarray().astype(np.float64)
    want = np.zeros((2, W, W))
    for a, c, s_ in zip(r0, c0, sl):
        blk = full[a : a + W, c : c + W].copy()
        ii, jj = np.arange(W)[:, None] + a, np.arange(W)[None, :] + c
        blk[(jj - ii) < 2] = 0.0
        want[s_] += blk
    for order in (np.arange(len(r0)), rng.permutation(len(r0))):
        acc = np.zeros(2 * nat.acc_stride(W))
        nv = nat.accumulate_region(0, nb, ip, col, cnt, None, None, None, np.ascontiguousarray(r0[order]),
                                   np.ascontiguousarray(c0[order]), np.ascontiguousarray(sl[order]), W, 2, 2, 0, acc)
        out = nat.acc_export(acc, W, 2)
        assert nv == len(r0)
        assert np.array_equal(out["sum"], want)  # exact: small integers
        assert np.array_equal(out["n"], [5096, 8])


@pytest.mark.parametrize("ignore_diags", [2, 0, -4])
@pytest.mark.parametrize("nb,W,dens", [(300, 21, 30), (900, 83, 300), (64, 5, 2), (2000, 11, 1)])
def test_upper_triangle_input_is_mirrored_on_device(nb, W, dens, ignore_diags):
    """pup_region_create_upper (cooler's stored upper triangle, incl. pixels that leave the region) gives the same
    accumulators as the symmetric-CSR entry point and as the oracle.  ignore_diags >= 0: the diagonal mask removes
    the lower triangle and nothing is mirrored; ignore_diags < 0: the lower triangle is mirrored in on the device."""
    nat = _cuda()
    from oracle.pileup_oracle import oracle_accumulate

    n_slots = 3
    ip, col, cnt, w, e, cov = random_region(nb, dens, seed=nb, nan_frac=0.05, with_expected=True)
    r0, c0, sl = random_windows(nb, W, 600, n_slots, seed=nb + 1, near_diag_frac=0.5)
    ref = oracle_accumulate(nb, ip, col, cnt, w, e, None, r0, c0, sl, W, ignore_diags, n_slots, ooe=True)
    # build the upper triangle the way a cooler holds it: rows of a LARGER matrix, so some columns are >= nb
    rows = np.repeat(np.arange(nb), np.diff(ip))
    up = col >= rows
    rng = np.random.default_rng(nb)
    extra_rows = rng.integers(0, nb, nb // 2)
    extra_cols = rng.integers(nb, nb + 500, nb // 2)
    ur = np.concatenate([rows[up], extra_rows])
    uc = np.concatenate([col[up], extra_cols])
    uv = np.concatenate([cnt[up], np.full(nb // 2, 7, dtype=np.int32)])
    o = np.lexsort((uc, ur))
    ur, uc, uv = ur[o], uc[o].astype(np.int32), uv[o].astype(np.int32)
    uip = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum(np.bincount(ur, minlength=nb), out=uip[1:])
    uip = uip.astype(np.int32)
    stride = nat.acc_stride(W)
    for memory in ("host", "device"):
        if memory == "host":
            reg = nat.Region(0, nb, uip, uc, uv, w, e, None, ignore_diags=ignore_diags, flags=nat.PUP_F_OOE, upper=True)
            acc = np.zeros(n_slots * stride)
            nv = reg.accumulate(r0, c0, sl, W, n_slots, 0, acc, want_n_valid=True)
        else:
            import torch

            dev = torch.device("cuda", 0)
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            tens = [t(x) for x in (uip, uc, uv, w, e, r0, c0, sl)]
            stream = torch.cuda.current_stream(dev).cuda_stream
            reg = nat.Region(0, nb, tens[0], tens[1], tens[2], tens[3], tens[4], None, ignore_diags=ignore_diags,
                             flags=nat.PUP_F_OOE, stream=stream, upper=True)
            acc = torch.zeros(n_slots * stride, dtype=torch.float64, device=dev)
            nv = reg.accumulate(tens[5], tens[6], tens[7], W, n_slots, 0, acc, stream=stream, want_n_valid=True)
        out = nat.acc_export(acc, W, n_slots)
        reg.close()
        _check_against_oracle(nv, out, ref)


@pytest.mark.parametrize("ignore_diags", [0, 2, 5, -1000000])
def test_ignore_diags_variants(ignore_diags):
    nat = _cuda()
    from oracle.pileup_oracle import oracle_accumulate

    nb, W, n_slots = 400, 11, 2
    ip, col, cnt, w, e, cov = random_region(nb, 40, seed=3, nan_frac=0.1, with_expected=True)
    r0, c0, sl = random_windows(nb, W, 800, n_slots, seed=5, near_diag_frac=0.7)
    ref = oracle_accumulate(nb, ip, col, cnt, w, e, None, r0, c0, sl, W, ignore_diags, n_slots, ooe=True)
    acc = np.zeros(n_slots * nat.acc_stride(W))
    nv = nat.accumulate_region(0, nb, ip, col, cnt, w, e, None, r0, c0, sl, W, ignore_diags, n_slots, 1, acc)
    _check_against_oracle(nv, nat.acc_export(acc, W, n_slots), ref)


def test_empty_and_degenerate_inputs():
    nat = _cuda()
    nb, W = 50, 5
    ip, col, cnt, w, e, cov = random_region(nb, 5, seed=1, nan_frac=0.0)
    acc = np.zeros(nat.acc_stride(W))
    z = np.zeros(0, dtype=np.int32)
    assert nat.accumulate_region(0, nb, ip, col, cnt, None, None, None, z, z, z, W, 2, 1, 0, acc) == 0
    assert not acc.any()
    # all windows out of bounds
    r0 = np.array([-1, 48, 10], dtype=np.int32)
    c0 = np.array([3, 3, 46], dtype=np.int32)
    sl = np.zeros(3, dtype=np.int32)
    assert nat.accumulate_region(0, nb, ip, col, cnt, None, None, None, r0, c0, sl, W, 2, 1, 0, acc) == 0
    assert not acc.any()
    # empty matrix
    ip0 = np.zeros(nb + 1, dtype=np.int32)
    r0 = np.array([0, 10], dtype=np.int32)
    c0 = np.array([20, 30], dtype=np.int32)
    sl = np.zeros(2, dtype=np.int32)
    nv = nat.accumulate_region(0, nb, ip0, z, z, None, None, None, r0, c0, sl, W, 2, 1, 0, acc)
    out = nat.acc_export(acc, W, 1)
    assert nv == 2 and out["n"][0] == 2 and not out["sum"].any() and (out["num"] == 2).all()
    # error reporting
    with pytest.raises(nat.NativeError):
        nat.accumulate_region(0, nb, ip, col, cnt, None, None, None, r0, c0, sl, W, 2, 1, 1, acc)  # OOE without expected


def test_linearity_and_permutation_large():
    """Size-independent properties on a matrix larger than L2: windows permuted / split across calls give the same
    accumulators, and duplicating every window doubles them."""
    nat = _cuda()
    import torch

    dev = torch.device("cuda", 0)
    from coolpuppy_b200.synthetic import synthetic_region

    nb, W, n_slots = 12000, 83, 4
    reg_t = synthetic_region(nb, depth=500.0, seed=11, device=dev, nan_frac=0.03)
    stream = torch.cuda.current_stream(dev).cuda_stream
    reg = nat.Region(0, nb, reg_t["indptr"], reg_t["col"], reg_t["count"], reg_t["weight"], reg_t["expected"], None,
                     ignore_diags=2, flags=nat.PUP_F_OOE, stream=stream)
    g = torch.Generator(device="cpu").manual_seed(5)
    n = 200_000
    r0 = torch.randint(0, nb - W, (n,), generator=g, dtype=torch.int32)
    c0 = torch.clamp(r0 + torch.randint(0, 3000, (n,), generator=g, dtype=torch.int32), max=nb - W)
    sl = torch.randint(0, n_slots, (n,), generator=g, dtype=torch.int32)
    r0, c0, sl = r0.to(dev), c0.to(dev), sl.to(dev)
    stride = nat.acc_stride(W)

    def run(parts):
        acc = torch.zeros(n_slots * stride, dtype=torch.float64, device=dev)
        for idx in parts:
            reg.accumulate(r0[idx].contiguous(), c0[idx].contiguous(), sl[idx].contiguous(), W, n_slots, 0, acc, stream=stream)
        return nat.acc_export(acc, W, n_slots, device=0, stream=stream)

    base = run([torch.arange(n, device=dev)])
    perm = torch.randperm(n, generator=g).to(dev)
    split = run([perm[: n // 3], perm[n // 3 :]])
    dup = run([torch.arange(n, device=dev), perm])
    assert base["n"].sum() == n
    assert np.array_equal(base["num"], split["num"]) and np.array_equal(base["n"], split["n"])
    np.testing.assert_allclose(split["sum"], base["sum"], rtol=1e-9)
    assert np.array_equal(2 * base["num"], dup["num"])
    np.testing.assert_allclose(dup["sum"], 2 * base["sum"], rtol=1e-9)
    # spot-check a sample of the same windows against the dense oracle
    from oracle.pileup_oracle import oracle_accumulate

    k = 300
    h = lambda t: t.cpu().numpy()
    ref = oracle_accumulate(nb, h(reg_t["indptr"]), h(reg_t["col"]), h(reg_t["count"]), h(reg_t["weight"]),
                            h(reg_t["expected"]), None, h(r0[:k]), h(c0[:k]), h(sl[:k]), W, 2, n_slots, ooe=True)
    acc = torch.zeros(n_slots * stride, dtype=torch.float64, device=dev)
    reg.accumulate(r0[:k].contiguous(), c0[:k].contiguous(), sl[:k].contiguous(), W, n_slots, 0, acc, stream=stream)
    out = nat.acc_export(acc, W, n_slots, device=0, stream=stream)
    _check_against_oracle(k, out, ref)
    reg.close()


def test_async_upload_pipeline_matches_oracle():
    """The end-to-end pattern of bench.py: regions created from pinned HOST upper triangles with PUP_F_ASYNC on a
    prepare stream, window arrays through pup_upload, pile-ups with PUP_F_ASYNC on a compute stream, everything
    adding into one device accumulator -- equals the sum of the oracle's per-region accumulators."""
    nat = _cuda()
    import torch

    from oracle.pileup_oracle import oracle_accumulate

    dev = torch.device("cuda", 0)
    W, n_slots = 21, 3
    stride = nat.acc_stride(W)
    regions = []
    ref_sum, ref_num, ref_n = 0, 0, 0
    for k, (nb, dens) in enumerate([(900, 60), (300, 30), (1500, 120), (64, 3)]):
        ip, col, cnt, w, e, cov = random_region(nb, dens, seed=100 + k, nan_frac=0.05, with_expected=True)
        r0, c0, sl = random_windows(nb, W, 700, n_slots, seed=200 + k)
        ref = oracle_accumulate(nb, ip, col, cnt, w, e, None, r0, c0, sl, W, 2, n_slots, ooe=True)
        ref_sum, ref_num, ref_n = ref_sum + np.nan_to_num(ref["sum"], posinf=np.inf), ref_num + ref["num"], ref_n + ref["n"]
        rows = np.repeat(np.arange(nb), np.diff(ip))
        up = col >= rows
        uip = np.zeros(nb + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows[up], minlength=nb), out=uip[1:])
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        regions.append(dict(nb=nb, uip=pin(uip.astype(np.int32)), uc=pin(col[up]), uv=pin(cnt[up]), w=pin(w), e=pin(e),
                            win=[pin(r0), pin(c0), pin(sl)]))
    acc = torch.zeros(n_slots * stride, dtype=torch.float64, device=dev)
    s_prep, s_comp = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    dwin = [[torch.empty_like(t, device=dev) for t in r["win"]] for r in regions]
    torch.cuda.synchronize()
    ASYNC = nat.PUP_F_ASYNC

    def upload(k):
        r = regions[k]
        reg = nat.Region(0, r["nb"], r["uip"], r["uc"], r["uv"], r["w"], r["e"], None, ignore_diags=2,
                         flags=nat.PUP_F_OOE | ASYNC, stream=s_prep.cuda_stream, upper=True)
        for d, h in zip(dwin[k], r["win"]):
            nat.upload(0, d, h, stream=s_prep.cuda_stream)
        return reg, s_prep.record_event()

    for _ in range(2):  # twice: the second pass re-uses pool memory and the landing buffers
        acc.zero_()
        torch.cuda.synchronize()
        nxt = upload(0)
        for k in range(len(regions)):
            reg, ready = nxt
            nxt = upload(k + 1) if k + 1 < len(regions) else None
            s_comp.wait_event(ready)
            reg.accumulate(dwin[k][0], dwin[k][1], dwin[k][2], W, n_slots, ASYNC, acc, stream=s_comp.cuda_stream)
            s_prep.wait_event(s_comp.record_event())
            with torch.cuda.stream(s_prep):
                reg.close()
        torch.cuda.synchronize()
        out = nat.acc_export(acc, W, n_slots, device=0)
        assert np.array_equal(out["n"], ref_n) and np.array_equal(out["num"], ref_num)
        m = np.isfinite(ref_sum)
        assert np.array_equal(np.isinf(out["sum"]), np.isinf(ref_sum))
        np.testing.assert_allclose(out["sum"][m], ref_sum[m], rtol=RTOL, atol=1e-300)


def test_legacy_loop_ref_statistically(fixtures_dir):
    """BASELINE.json north_star names tests/loop_ref.np.txt: a Monte-Carlo output of a pre-1.0 CLI (random control
    shifts of another generator), so the GPU path with the options of its header (coverage_norm,
    nshifts 10, seed 0, unbalanced, mindist 0, pad 100 kb) is compared statistically.  (tests/bed2_ref.np.txt needs the
    removed `bed2` option, whose semantics are not recoverable from the 1.1.0 tree: an all-vs-all (+, -) emulation
    correlates at 0.07 with it -- see DESIGN.md section 2.)"""
    _cuda()
    import os

    import pandas as pd

    from coolpuppy_b200 import coolpup as cp
    from coolpuppy_b200.coolio import Cooler

    clr = Cooler(os.path.join(fixtures_dir, "Scc1-control.10000.cool"))
    loops = pd.read_csv(os.path.join(fixtures_dir, "CH12_loops_Rao.bed"), sep="\t", header=None).iloc[:, :6]
    loops.columns = ["chrom1", "start1", "end1", "chrom2", "start2", "end2"]
    ref = np.loadtxt(os.path.join(fixtures_dir, "loop_ref.np.txt"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pups = cp.pileup(clr, loops, features_format="bedpe", clr_weight_name=None, flank=100_000, mindist=0, nshifts=10,
                         seed=0, coverage_norm=True)  # the coverage columns are computed in memory (expected.coverage)
    mine = np.asarray(pups["data"].iloc[0], dtype=float)
    assert np.corrcoef(mine.ravel(), ref.ravel())[0, 1] > 0.9
    assert abs(mine[10, 10] / ref[10, 10] - 1) < 0.15
    assert np.median(np.abs(mine - ref) / ref) < 0.06


def test_legacy_tad_ref_statistically(fixtures_dir):
    """tests/tad_ref.np.txt: the legacy (pre-1.0 CLI) golden of a local, rescaled TAD pile-up -- rescale_pad 1, rescale_size
    99, nshifts 10, seed 0, coverage_norm, unbalanced (header of the file).  Like loop_ref it is a Monte-Carlo output of
    another random generator, and the legacy code did not mask the first diagonals (it has no NaN cells; the current
    reference, and this path, leave the cells the masked diagonals touch NaN), so it is compared statistically on the
    cells both have: a flat matrix around 1 -- median relative difference 4 %, Pearson 0.77."""
    _cuda()
    import os

    import pandas as pd

    from coolpuppy_b200 import coolpup as cp
    from coolpuppy_b200.coolio import Cooler

    clr = Cooler(os.path.join(fixtures_dir, "Scc1-control.10000.cool"))
    tads = pd.read_csv(os.path.join(fixtures_dir, "CH12_TADs_Rao.bed"), sep="\t", header=None).iloc[:, :3]
    tads.columns = ["chrom", "start", "end"]
    ref = np.loadtxt(os.path.join(fixtures_dir, "tad_ref.np.txt"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pups = cp.pileup(clr, tads, features_format="bed", local=True, rescale=True, rescale_flank=1.0, rescale_size=99,
                         clr_weight_name=None, nshifts=10, seed=0, coverage_norm=True)
    mine = np.asarray(pups["data"].iloc[0], dtype=float)
    assert mine.shape == ref.shape == (99, 99)
    m = np.isfinite(mine) & np.isfinite(ref)
    assert m.sum() > 0.95 * m.size
    assert np.corrcoef(mine[m], ref[m])[0, 1] > 0.6
    assert np.median(np.abs(mine[m] - ref[m]) / np.abs(ref[m])) < 0.08


@pytest.mark.parametrize("mode", ["raw", "balanced", "balanced_ooe", "balanced_expblocks", "raw_coverage", "raw_local"])
@pytest.mark.parametrize("nb,rs,hmax,nwin,n_slots", [(300, 9, 40, 300, 3), (500, 11, 8, 400, 2), (900, 99, 260, 60, 2),
                                                      (400, 7, 90, 200, 4)])
def test_rescale_kernel_vs_oracle_random(mode, nb, rs, hmax, nwin, n_slots):
    """pup_accumulate_rescaled on seeded random CSR + windows of random sizes (smaller and larger than rescale_size,
    empty, outside the region, on the diagonal) == the per-window restatement with the real scipy zoom."""
    nat = _cuda()
    from oracle.pileup_oracle import oracle_accumulate_rescaled

    ip, col, cnt, w, e, cov = random_region(nb, 60, seed=nb + rs, nan_frac=0.05, with_expected=True, with_cov=True)
    e = np.where(np.isnan(e) | (e == 0), 1.0, e) if mode == "balanced_ooe" else e  # x / 0 = inf: the zoom would spread it
    rng = np.random.default_rng(5 * nb + rs)
    local = mode == "raw_local"
    h = rng.integers(0 if hmax > 20 else 1, hmax, nwin).astype(np.int32)
    wd = h.copy() if local else rng.integers(1, hmax, nwin).astype(np.int32)
    r0 = rng.integers(-3, nb - 2, nwin).astype(np.int32)
    c0 = r0.copy() if local else np.clip(r0 + rng.integers(-hmax, 3 * hmax, nwin), -2, nb).astype(np.int32)
    sl = np.sort(rng.integers(0, n_slots, nwin)).astype(np.int32)
    md = (rng.random(nwin) < 0.4).astype(np.int32) if mode == "balanced_expblocks" else None
    weight = None if mode.startswith("raw") else w
    expected = e if mode in ("balanced_ooe", "balanced_expblocks") else None
    if md is not None:
        expected = np.where(np.isnan(e), 0.5, e)
    coverage = cov if mode == "raw_coverage" else None
    ref = oracle_accumulate_rescaled(nb, ip, col, cnt, weight, expected, coverage, r0, c0, h, wd, sl, md, rs, 2, n_slots,
                                     ooe=mode == "balanced_ooe", local=local)
    import torch

    dev = torch.device("cuda", 0)
    flags = (nat.PUP_F_OOE if mode == "balanced_ooe" else 0)
    reg = nat.Region(0, nb, ip, col, cnt, weight, expected, coverage, ignore_diags=2, flags=flags)
    acc = torch.zeros(n_slots * nat.acc_stride(rs), dtype=torch.float64, device=dev)
    aflags = (nat.PUP_F_COVERAGE if coverage is not None else 0) | (nat.PUP_F_LOCAL if local else 0)
    nv = reg.accumulate_rescaled(r0, c0, h, wd, sl, md, rs, n_slots, aflags, acc, want_n_valid=True)
    out = nat.acc_export(acc, rs, n_slots, device=0, want_cov=coverage is not None)
    reg.close()
    assert nv == int(ref["n"].sum())
    assert np.array_equal(out["n"], ref["n"])
    assert np.array_equal(out["num"], ref["num"])
    np.testing.assert_allclose(out["sum"], ref["sum"], rtol=RTOL, atol=1e-300)
    if coverage is not None:
        np.testing.assert_allclose(out["cov_start"], ref["cov_start"], rtol=RTOL)
        np.testing.assert_allclose(out["cov_end"], ref["cov_end"], rtol=RTOL)


@pytest.mark.parametrize("mode", ["raw", "balanced", "balanced_ooe", "balanced_expctrl"])
@pytest.mark.parametrize("nb,W,dens,nwin,n_slots", [(1500, 83, 800, 3000, 3), (600, 21, 300, 5000, 4), (1200, 45, 500, 2000, 2),
                                                      (1400, 89, 900, 600, 2), (900, 64, 10, 1500, 2), (1600, 203, 900, 250, 2), (1300, 121, 700, 300, 3)])
def test_dense_band_path_vs_oracle(monkeypatch, mode, nb, W, dens, nwin, n_slots):
    """Windows inside the dense diagonal band go through k_pileup_dense (register tiles over a dense copy of the
    matrix), the others through the sparse kernel; a budget / density setting that puts most windows into the band, and
    the default one, both give the oracle's accumulators -- and the same as with the band switched off."""
    nat = _cuda()
    from oracle.pileup_oracle import oracle_accumulate

    cfg = MODES[mode]
    ip, col, cnt, w, e, cov = random_region(nb, dens, seed=nb + W, nan_frac=0.04, with_expected=True)
    weight = w if cfg["bal"] else None
    expected = e if cfg.get("exp") else None
    rng = np.random.default_rng(nb * 7 + W)
    r0 = rng.integers(0, nb - W, nwin).astype(np.int32)
    c0 = np.clip(r0 + rng.integers(-W, nb // 2, nwin), 0, nb - W).astype(np.int32)  # most windows near the diagonal
    sl = rng.integers(0, n_slots, nwin).astype(np.int32)
    ref = oracle_accumulate(nb, ip, col, cnt, weight, expected, None, r0, c0, sl, W, 2, n_slots,
                            ooe=cfg["flags"] == 1, expctrl=cfg["flags"] == 2)
    outs = []
    for env in ({"PUP_BAND_PCT": "3000", "PUP_BAND_DENSITY_PCT": "1"}, {}, {"PUP_BAND": "0"}):
        with monkeypatch.context() as mp:
            for k, v in env.items():
                mp.setenv(k, v)
            acc = np.zeros(n_slots * nat.acc_stride(W))
            nv = nat.accumulate_region(0, nb, ip, col, cnt, weight, expected, None, r0, c0, sl, W, 2, n_slots, cfg["flags"], acc)
            out = nat.acc_export(acc, W, n_slots, want_expected=True)
        _check_against_oracle(nv, out, ref)
        outs.append(out)
    for o in outs[:2]:
        assert np.array_equal(o["num"], outs[2]["num"]) and np.array_equal(o["n"], outs[2]["n"])

"""GPU tests of the device-side window generation: the MT19937 control-shift replay (``pup_control_shifts``) against
numpy's own stream, and ``pup_pair_windows_device`` against the host window builder that is itself pinned to the
reference's recorded ``pos_stream`` (tests/test_host_pipeline.py::test_window_arrays_match_reference_stream)."""
import warnings

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _cuda():
    from coolpuppy_b200 import _native

    if _native.device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return _native


@pytest.mark.parametrize("seed,burn", [(0, 0), (1, 7), (12345, 623), (7, 624 * 3 + 5)])
def test_control_shifts_replay_numpy_stream(seed, burn):
    nat = _cuda()
    import torch

    from coolpuppy_b200._coords import _draw_shifts

    segs = np.array([5, 1, 700, 33, 1249, 2, 20000, 624, 623, 1], dtype=np.int64)
    np.random.seed(seed)
    if burn:
        np.random.randint(0, 2**31 - 1, burn)
    state0 = np.random.get_state()
    want = np.concatenate([_draw_shifts(int(n), 100_000, 1_000_000, 10_000) for n in segs])
    after_host = np.random.get_state()
    tail_host = np.random.random(5)
    # device replay from the same starting state
    np.random.set_state(state0)
    rng = nat.DeviceRng(0)
    dbin = torch.empty(int(segs.sum()), dtype=torch.int32, device="cuda:0")
    rng.control_shifts(segs[:4], 100_000, 1_000_000, 10_000, dbin[: int(segs[:4].sum())])
    rng.control_shifts(segs[4:], 100_000, 1_000_000, 10_000, dbin[int(segs[:4].sum()) :])  # state persists between calls
    torch.cuda.synchronize()
    got = dbin.cpu().numpy().astype(np.int64)
    assert np.array_equal(got, want)
    rng.store()
    rng.close()
    st = np.random.get_state()
    assert st[2] == after_host[2] and np.array_equal(st[1], after_host[1])
    assert np.array_equal(np.random.random(5), tail_host)
    # advance-only mode (another rank's region) ends in the same state
    np.random.set_state(state0)
    rng = nat.DeviceRng(0)
    rng.control_shifts(segs, 100_000, 1_000_000, 10_000, None)
    rng.store()
    rng.close()
    assert np.array_equal(np.random.random(5), tail_host)


def test_control_shifts_many_small_segments():
    """Thousands of short segments: every alignment of a segment's end with the 624-word MT19937 blocks occurs,
    including a segment whose last accepted word is followed by rejected words inside the same block (those words
    belong to the sign draws -- a round-2 bug let them vanish)."""
    nat = _cuda()
    import torch

    from coolpuppy_b200._coords import _draw_shifts

    rs = np.random.RandomState(5)
    segs = rs.randint(1, 700, 4000).astype(np.int64)
    np.random.seed(11)
    state0 = np.random.get_state()
    want = np.concatenate([_draw_shifts(int(n), 100_000, 1_000_000, 10_000) for n in segs])
    tail_host = np.random.random(4)
    np.random.set_state(state0)
    rng = nat.DeviceRng(0)
    dbin = torch.empty(int(segs.sum()), dtype=torch.int32, device="cuda:0")
    rng.control_shifts(segs, 100_000, 1_000_000, 10_000, dbin)
    torch.cuda.synchronize()
    got = dbin.cpu().numpy().astype(np.int64)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, (len(bad), bad[:10])
    rng.store()
    rng.close()
    assert np.array_equal(np.random.random(4), tail_host)


def test_control_shifts_other_ranges():
    nat = _cuda()
    import torch

    from coolpuppy_b200._coords import _draw_shifts

    for lo, hi, res in [(10, 11_000, 1000), (5, 7, 1), (0, 2**20 + 3, 4096), (1, 2**31 - 1, 1_000_000)]:
        np.random.seed(lo + hi)
        s0 = np.random.get_state()
        want = _draw_shifts(3000, lo, hi, res)
        np.random.set_state(s0)
        rng = nat.DeviceRng(0)
        dbin = torch.empty(3000, dtype=torch.int32, device="cuda:0")
        rng.control_shifts([3000], lo, hi, res, dbin)
        torch.cuda.synchronize()
        rng.close()
        assert np.array_equal(dbin.cpu().numpy().astype(np.int64), want), (lo, hi, res)


PAIR_CASES = ["toy_controls", "toy_strand_dist_ctrl", "toy_dist_edges", "toy_bywindow", "toy_flipneg_igo", "toy_strand_igo",
              "scc1_ctcf_pairs_strand_dist", "scc1_ctcf_pairs_arms", "scc1_ctcf_pairs_flip_ooe", "toy_mindist_auto"]


def _pileupper(name):
    from coolpuppy_b200 import coolpup as cp

    clr, feats, kw = gu.case_inputs(name)
    kw = dict(kw)
    seed, nshifts = kw.get("seed"), kw.get("nshifts", 0)
    view = kw.get("view_df")
    view = cp.make_cooler_view(clr) if view is None else view
    if seed is not None:
        np.random.seed(seed)
    cc = cp.CoordCreator(feats, clr.binsize, features_format="bed", flank=kw.get("flank", 100000),
                         chroms=list(view["chrom"].unique()), nshifts=nshifts, mindist=kw.get("mindist", "auto"),
                         maxdist=kw.get("maxdist"), seed=seed)
    exp = kw.get("expected_df")
    pu = cp.PileUpper(clr, cc, view_df=view, clr_weight_name=kw.get("clr_weight_name", "weight"), expected=exp if exp is not None else False,
                      expected_value_col=kw.get("expected_value_col", "balanced.avg"), ooe=kw.get("ooe", True), control=nshifts > 0,
                      flip_negative_strand=kw.get("flip_negative_strand", False), ignore_diags=kw.get("min_diag", 2))
    groupby, modify, post = [], None, None
    if kw.get("by_strand"):
        groupby += ["strand1", "strand2"]
    if kw.get("by_distance") is not None and kw.get("by_distance") is not False:
        from functools import partial

        bd = kw["by_distance"]
        edges = "default" if bd is True else [int(x) for x in bd]
        modify = partial(cp.bin_distance_intervals, band_edges=pu._distance_edges(edges))
        groupby += ["distance_band"]
    if kw.get("by_window"):
        post = cp.group_by_region
    return pu, groupby, modify, post, kw


@pytest.mark.parametrize("name", PAIR_CASES)
@pytest.mark.parametrize("parts", [1, 3])
def test_device_windows_equal_host_windows(name, parts):
    """r0 / c0 / slot arrays generated on the GPU == the host builder's arrays, element by element, in emission
    order (incl. the seeded control shifts), for whole regions and for row-anchor ranges of them (how a region's
    windows are split over ranks); the first-appearance table gives the same group order."""
    nat = _cuda()
    import torch

    pu, groupby, modify, post, kw = _pileupper(name)
    igo = kw.get("ignore_group_order", False)
    plan = pu._plan(groupby, igo, modify, post)
    plan["band_edges"] = pu._band_edges(plan)
    assert pu._device_windows_ok(plan)
    seed = kw.get("seed")
    if seed is not None:
        np.random.seed(seed)
    state0 = np.random.get_state()
    host = pu._prepare(plan, None, None)
    state_host = np.random.get_state()
    np.random.set_state(state0)
    job = pu._prepare_device(plan, None, None)
    assert job["n_slots"] == host["n_slots"] and job["n_keys"] == host["n_keys"]
    dev = torch.device("cuda", 0)
    rng = nat.DeviceRng(0)
    first = torch.full((job["n_keys"],), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
    by_name = {b["name"]: b for b in host["built"]}
    edges = plan["band_edges"]
    edges = None if edges is None else np.ascontiguousarray(edges, dtype=np.float64)
    for it in job["items"]():
        dbin = None
        if len(it["segs"]):
            dbin = torch.empty(int(it["total"]) * it["nctrl"], dtype=torch.int32, device=dev)
            rng.control_shifts(it["segs"], pu.CC.minshift, pu.CC.maxshift, pu.resolution, dbin)
        if not it["owned"]:
            assert it["name"] not in by_name
            continue
        b = by_name[it["name"]]
        targets = 2 if it["ident"] is not None else 1
        n_all = int(it["total"]) * (1 + it["nctrl"])
        m = len(it["center"])
        cuts = np.linspace(0, m, parts + 1).astype(int)  # row-anchor ranges [k_lo, k_hi)
        idx1 = np.repeat(np.asarray(b["rw"].idx1), targets)
        for part in range(parts):
            k_lo, k_hi = int(cuts[part]), int(cuts[part + 1])
            q_part, tot_part = nat.pair_windows_count(it["center"], pu.CC.mindist, pu.CC.maxdist, k_lo, k_hi)
            n_mine = tot_part * (1 + it["nctrl"])
            idx = np.nonzero((idx1 >= k_lo) & (idx1 < k_hi))[0]
            assert len(idx) == n_mine * targets
            outs = tuple(torch.full((n_mine * targets,), -7, dtype=torch.int32, device=dev) for _ in range(3))
            nat.pair_windows_device(0, it["stbin"], it["center"], pu.CC.mindist, pu.CC.maxdist, it["nctrl"], it["per_offset"], dbin,
                                    it["nb"], job["W"], it["key1"], it["key2"], edges, it["band_weight"],
                                    0 if it["flipval"] is None else (1 if pu.flip_negative_strand else 2),
                                    bool(plan["flip"] and plan["ignore_group_order"]), it["flipval"], it["ident"], job["nk"],
                                    job["nf"], k_lo, k_hi, q_part, it["index"], outs[0], outs[1], outs[2],
                                    first_seen=first if parts == 1 else None)
            torch.cuda.synchronize()
            for got, want in zip(outs, (b["w_r0"], b["w_c0"], b["slot"])):
                assert np.array_equal(got.cpu().numpy().astype(np.int64), np.asarray(want)[idx].astype(np.int64))
    rng.store()
    rng.close()
    st = np.random.get_state()
    assert st[2] == state_host[2] and np.array_equal(st[1], state_host[1])
    if parts == 1:
        fv = first.cpu().numpy()
        got = {int(k): (int(fv[k] >> 62), int((fv[k] >> 40) & 0xFFFFF), int(fv[k] & ((1 << 40) - 1)))
               for k in np.nonzero(fv != np.iinfo(np.int64).max)[0]}
        assert got == {int(k): tuple(int(x) for x in v) for k, v in host["first"].items()}


@pytest.mark.parametrize("name", PAIR_CASES + ["toy_zero_expected_strand", "toy_zero_expected_bywindow", "toy_strand_notooe",
                                               "toy_strand_rawcov"])
@pytest.mark.parametrize("device_windows", ["0", "1"])
def test_golden_cases_on_both_window_paths(monkeypatch, name, device_windows):
    """The bed-pair golden cases through pileup() with windows laid out by the host builder and by the GPU."""
    _cuda()
    from coolpuppy_b200 import coolpup as cp
    from test_gpu_parity import _compare_rows

    monkeypatch.setenv("PUP_DEVICE_WINDOWS", device_windows)
    clr, feats, kw = gu.case_inputs(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pups = cp.pileup(clr, feats, **kw)
    _compare_rows(pups, name)
    # the global random stream is left exactly where the reference's serial run leaves it
    if kw.get("nshifts", 0) > 0:
        after = np.random.random(3)
        monkeypatch.setenv("PUP_DEVICE_WINDOWS", "0")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cp.pileup(clr, feats, **kw)
        assert np.array_equal(np.random.random(3), after)


def test_device_windows_at_bench_scale():
    """The whole configs[3] genome (6.4 k sites, 1.1e7 windows, 24 chromosomes, up to 510 sites per chromosome --
    several thread tiles per pair offset, thousands of MT19937 segments): device windows == host windows."""
    nat = _cuda()
    import torch

    from coolpuppy_b200 import coolpup as cp
    from coolpuppy_b200.coolio import ChromCooler
    from coolpuppy_b200.synthetic import HG38, synthetic_sites

    sizes = dict(HG38)
    sites, n_pairs = synthetic_sites(1_000_000, chromsizes=sizes, binsize=10_000, flank=410_000, seed=1237)
    nbins = sum(-(-L // 10_000) for L in sizes.values())
    clr = ChromCooler(sizes, 10_000, {}, {"weight": np.ones(nbins)})
    np.random.seed(0)
    cc = cp.CoordCreator(sites, 10_000, features_format="bed", flank=410_000, nshifts=10, mindist="auto", seed=0)
    pu = cp.PileUpper(clr, cc, clr_weight_name="weight", control=True)
    plan = pu._plan([], False, None, None)
    plan["band_edges"] = None
    state0 = np.random.get_state()
    host = pu._prepare(plan, None, None)
    state_host = np.random.get_state()
    np.random.set_state(state0)
    job = pu._prepare_device(plan, None, None)
    dev = torch.device("cuda", 0)
    rng = nat.DeviceRng(0)
    by_name = {b["name"]: b for b in host["built"]}
    for it in job["items"]():
        dbin = None
        if len(it["segs"]):
            dbin = torch.empty(int(it["total"]) * it["nctrl"], dtype=torch.int32, device=dev)
            rng.control_shifts(it["segs"], pu.CC.minshift, pu.CC.maxshift, pu.resolution, dbin)
        if not it["owned"]:
            continue
        b = by_name[it["name"]]
        n_all = int(it["total"]) * (1 + it["nctrl"])
        outs = tuple(torch.full((n_all,), -7, dtype=torch.int32, device=dev) for _ in range(3))
        nat.pair_windows_device(0, it["stbin"], it["center"], pu.CC.mindist, pu.CC.maxdist, it["nctrl"], it["per_offset"], dbin,
                                it["nb"], job["W"], None, None, None, 0, 0, False, None, None, job["nk"], job["nf"], 0,
                                len(it["center"]), None, it["index"], outs[0], outs[1], outs[2])
        torch.cuda.synchronize()
        for got, want in zip(outs, (b["w_r0"], b["w_c0"], b["slot"])):
            assert np.array_equal(got.cpu().numpy().astype(np.int64), np.asarray(want).astype(np.int64)), it["name"]
    rng.store()
    rng.close()
    st = np.random.get_state()
    assert st[2] == state_host[2] and np.array_equal(st[1], state_host[1])

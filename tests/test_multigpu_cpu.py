"""world_size-2 gloo test of the multi-GPU path's host logic: region sharding (LPT), group-dictionary merge,
all-reduce of the accumulators -- with the emulated kernels standing in for the GPUs."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_assign_balances():
    from coolpuppy_b200.multigpu import lpt_assign

    costs = [100, 90, 50, 40, 30, 20, 10, 5]
    owner = lpt_assign(costs, 3)
    loads = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(3)]
    assert max(loads) - min(loads) <= 20 and sorted(set(owner)) == [0, 1, 2]
    assert lpt_assign([], 4) == []


WORKER = textwrap.dedent(
    """
    import os, sys, warnings, json
    sys.path[:0] = [{root!r}, os.path.join({root!r}, "tests"), os.path.join({root!r}, "tests", "golden")]
    import numpy as np, pytest, torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size={world})
    import emulator, golden_util as gu
    from coolpuppy_b200 import coolpup as cp
    from coolpuppy_b200.multigpu import RegionSharder
    from oracle.pileup_oracle import key_repr
    mp = pytest.MonkeyPatch(); emulator.install(mp)
    sharder = RegionSharder()
    ok = True
    for name in {cases!r}:
        clr, feats, kw = gu.case_inputs(name)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pups = cp.pileup(clr, feats, dist=sharder, **kw)
        z, _ = gu.load_golden(name)
        if "group" in pups.columns:
            keys = [key_repr(g) for g in pups["group"]]
        else:
            keys = [repr((r.chrom, int(r.start), int(r.end))) for r in pups.itertuples()]
        ok &= keys == [str(k) for k in z["row_keys"]]
        for i in range(len(keys)):
            b = z[f"row{{i}}.data"]; a = np.asarray(pups["data"].iloc[i], dtype=float)
            m = np.isfinite(b)
            ok &= np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[m], b[m], rtol=1e-9)
            ok &= int(pups["n"].iloc[i]) == int(z[f"row{{i}}.n"]) and np.array_equal(np.asarray(pups["num"].iloc[i]), z[f"row{{i}}.num"])
            if f"row{{i}}.vertical_stripe" in z.files:
                for f in ("vertical_stripe", "horizontal_stripe"):
                    ok &= np.allclose(np.asarray(pups[f].iloc[i], dtype=float), z[f"row{{i}}.{{f}}"], rtol=1e-9, equal_nan=True)
                ok &= np.array_equal(np.asarray(pups["coordinates"].iloc[i]).astype(str), z[f"row{{i}}.coordinates"])
        if not ok:
            print("FAILED", name); break
    print("RESULT", sharder.rank, ok)
    dist.destroy_process_group()
    """
)


def _run_workers(tmp_path, world, cases, port_base):
    port = port_base + os.getpid() % 2000
    script = tmp_path / f"worker{world}.py"
    script.write_text(WORKER.format(root=ROOT, port=port, cases=cases, world=world))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"RESULT {r} True" in o, o


def test_two_rank_gloo_pileup_matches_golden(tmp_path):
    """pileup(dist=...) on 2 ranks: every toy region is cut into two strided window parts (one per rank), the group
    dictionary / first-appearance order is merged over the ranks, one all-reduce merges the accumulators; stripes
    and the +inf merge quirk are gathered."""
    cases = ["toy_strand_ooe", "toy_strand_dist_ctrl", "scc1_loops_dist", "scc1_ctcf_pairs_arms", "toy_bywindow",
             "toy_zero_expected_strand", "toy_zero_expected_bywindow", "toy_stripes_strand_ooe", "toy_flipneg_igo"]
    _run_workers(tmp_path, 2, cases, 29500)


def test_more_ranks_than_regions(tmp_path):
    """4 ranks, 2 view regions: ranks without any window must neither crash nor block the collectives."""
    cases = ["toy_strand_ooe", "toy_bywindow", "toy_controls", "toy_stripes", "toy_zero_expected_ooe"]
    _run_workers(tmp_path, 4, cases, 33500)


def test_split_heavy_units_cover_every_window_once_and_balance():
    """bench.py's N-GPU sharding: heavy regions are cut by windows; every window belongs to exactly one unit."""
    import numpy as np

    from coolpuppy_b200.multigpu import part_index, split_heavy
    from coolpuppy_b200.synthetic import HG38

    nb = np.array([-(-v // 10_000) for v in HG38.values()], dtype=np.float64)
    nwin = np.round(1.1e7 * nb**2 / (nb**2).sum()).astype(int)  # all-vs-all pairs ~ n^2
    cost = list(nwin * nb / 1e4)
    for world in (1, 2, 4, 8):
        units, ucost, owner = split_heavy(cost, world)
        assert len(units) == len(ucost) == len(owner)
        if world == 1:
            assert units == [(i, 0, 1) for i in range(len(cost))]
        covered = [np.zeros(n, dtype=int) for n in nwin]
        for (i, part, parts), o in zip(units, owner):
            assert 0 <= o < world and 0 <= part < parts <= world
            covered[i][part_index(int(nwin[i]), part, parts)] += 1
        assert all((c == 1).all() for c in covered)
        load = np.bincount(owner, weights=ucost, minlength=world)
        assert load.max() <= 1.03 * sum(cost) / world  # whole chromosomes only: 1.17 at 8 ranks
        assert abs(sum(ucost) - sum(cost)) < 1e-6 * sum(cost)


SPLIT_WORKER = textwrap.dedent(
    """
    import os, sys
    sys.path[:0] = [{root!r}, os.path.join({root!r}, "tests")]
    import numpy as np, torch
    import torch.distributed as dist
    rank = int(sys.argv[1])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
    import emulator
    from synth import random_region, random_windows
    from coolpuppy_b200.multigpu import part_index, split_heavy
    from oracle.pileup_oracle import oracle_accumulate
    W, n_slots = 11, 3
    stride = emulator.layout(W)["stride"]
    regs, wins, ref = [], [], None
    for k, (nb, dens, nwin) in enumerate([(400, 40, 900), (120, 10, 60), (90, 8, 40)]):
        ip, col, cnt, w, e, cov = random_region(nb, dens, seed=k, nan_frac=0.05, with_expected=True)
        r0, c0, sl = random_windows(nb, W, nwin, n_slots, seed=10 + k)
        regs.append(emulator.EmuRegion(0, nb, ip, col, cnt, w, e, None, ignore_diags=2, flags=1))
        wins.append((r0, c0, sl))
        o = oracle_accumulate(nb, ip, col, cnt, w, e, None, r0, c0, sl, W, 2, n_slots, ooe=True)
        ref = o if ref is None else {{f: ref[f] + o[f] for f in ("sum", "num", "n")}}
    units, ucost, owner = split_heavy([len(w[0]) * r.nb for w, r in zip(wins, regs)], 2)
    assert any(parts > 1 for _, _, parts in units)          # the big region is cut by windows
    acc = torch.zeros(n_slots * stride, dtype=torch.float64)
    for (i, part, parts), o in zip(units, owner):
        if o != rank:
            continue
        sel = part_index(len(wins[i][0]), part, parts)
        regs[i].accumulate(wins[i][0][sel], wins[i][1][sel], wins[i][2][sel], W, n_slots, 0, acc)
    dist.all_reduce(acc)
    out = emulator.emu_export(acc, W, n_slots)
    ok = np.array_equal(out["n"], ref["n"]) and np.array_equal(out["num"], ref["num"])
    m = np.isfinite(ref["sum"])
    ok &= np.allclose(out["sum"][m], ref["sum"][m], rtol=1e-9)
    print("RESULT", rank, bool(ok))
    dist.destroy_process_group()
    """
)


def test_two_rank_gloo_window_split_units(tmp_path):
    """A region's window list split across two ranks (matrix replicated) + one all-reduce == all windows on one rank."""
    port = 31500 + os.getpid() % 2000
    script = tmp_path / "split_worker.py"
    script.write_text(SPLIT_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"RESULT {r} True" in o, o

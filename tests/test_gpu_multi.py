"""Multi-GPU product path on hardware: `pileup(dist=RegionSharder())` on 2 ranks over NCCL against the reference's
golden vectors (skipped when fewer than 2 CUDA devices are visible; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import os, sys, warnings
    sys.path[:0] = [{root!r}, os.path.join({root!r}, "tests"), os.path.join({root!r}, "tests", "golden")]
    import numpy as np, torch
    import torch.distributed as dist
    rank = int(sys.argv[1])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2,
                            device_id=torch.device("cuda", rank))
    import golden_util as gu
    from coolpuppy_b200 import coolpup as cp
    from coolpuppy_b200.multigpu import RegionSharder
    from oracle.pileup_oracle import key_repr
    sharder = RegionSharder()
    ok = True
    for name in {cases!r}:
        clr, feats, kw = gu.case_inputs(name)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pups = cp.pileup(clr, feats, dist=sharder, device=rank, **kw)
        z, _ = gu.load_golden(name)
        if "group" in pups.columns:
            keys = [key_repr(g) for g in pups["group"]]
        else:
            keys = [repr((r.chrom, int(r.start), int(r.end))) for r in pups.itertuples()]
        good = keys == [str(k) for k in z["row_keys"]]
        for i in range(len(keys) if good else 0):
            b = z[f"row{{i}}.data"]; a = np.asarray(pups["data"].iloc[i], dtype=float)
            m = np.isfinite(b)
            good &= np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[m], b[m], rtol=1e-6)
            good &= int(pups["n"].iloc[i]) == int(z[f"row{{i}}.n"]) and np.array_equal(np.asarray(pups["num"].iloc[i]), z[f"row{{i}}.num"])
        if not good:
            print("FAILED", name)
        ok &= bool(good)
    print("RESULT", rank, ok, cp._LAST_STATS.get("device_windows"))
    dist.destroy_process_group()
    """
)


def test_two_gpu_nccl_pileup_matches_golden(tmp_path):
    from coolpuppy_b200 import _native

    if _native.device_count() < 2:
        pytest.skip("needs 2 CUDA devices (gpurun --gpus 2)")
    cases = ["toy_controls", "toy_strand_dist_ctrl", "scc1_loops_ctrl", "scc1_ctcf_pairs_strand_dist", "scc1_ctcf_pairs_arms",
             "toy_bywindow", "toy_zero_expected_strand", "toy_stripes_strand_ooe", "toy_trans_ctrl", "scc1_ctcf_local_ooe"]
    port = 30500 + os.getpid() % 2000
    script = tmp_path / "nccl_worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port, cases=cases))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"RESULT {r} True" in o, o

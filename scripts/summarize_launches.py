"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (markdown + trimmed csv).

    python scripts/summarize_launches.py gpurun_out/launches_r1.csv profiles/r1_launches.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"void ", "", name)
    pre = "torch " if "at_cuda_detail" in name else ""
    if "DeviceRadixSort" in name:
        return pre + "cub::DeviceRadixSort*"
    if "DeviceScan" in name:
        return pre + "cub::DeviceScan*"
    return re.sub(r"\(.*", "", name)


def main(src, dst=None):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    out = [(int(r[0]), short(r[4]), float(r[-1]) / 1e6) for r in rows]  # id, kernel, ms
    if dst:
        with open(dst, "w") as f:
            f.write("launch_id,kernel,duration_ms\n")
            for i, k, ms in out:
                f.write(f"{i},{k},{ms:.6f}\n")
    # phases: split at the first k_window_keys (start of the first pup_accumulate)
    first_acc = next((n for n, (_, k, _) in enumerate(out) if k == "k_window_keys"), len(out))
    for title, part in (("before the first pup_accumulate (synthetic genome generation by torch, region preparation, "
                         "algorithmic-byte counting)", out[:first_acc]), ("from the first pup_accumulate on", out[first_acc:])):
        tot = OrderedDict()
        for _, k, ms in part:
            t = tot.setdefault(k, [0.0, 0])
            t[0] += ms
            t[1] += 1
        total = sum(v[0] for v in tot.values())
        print(f"## {title}: {total:.2f} ms total")
        for k, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            print(f"    {ms:10.3f} ms {n:6d} launches {100 * ms / total:6.1f}%  {k}")


if __name__ == "__main__":
    main(*sys.argv[1:])

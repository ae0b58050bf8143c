"""Build the CUDA library in-tree: ``python -m coolpuppy_b200.build``.

Produces ``coolpuppy_b200/libpileup_b200.so`` (sm_100a only, ``-lineinfo`` so ncu's source page maps to the
``.cu`` file).  The ``.so`` is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc", "pileup_b200.cu")
INC = os.path.join(os.path.dirname(PKG), "include")
OUT = os.path.join(PKG, "libpileup_b200.so")


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build():
    if not os.path.exists(OUT):
        return True
    newest = max(os.path.getmtime(SRC), os.path.getmtime(os.path.join(INC, "pileup_b200.h")))
    return os.path.getmtime(OUT) < newest


def build_native(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-shared", "-Xcompiler", "-fPIC", "-I", INC, "-o", OUT, SRC,
    ]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Builds libconsenrich_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m consenrich_b200.build [--force]

The library is plain C ABI (include/consenrich_b200.h); nothing here depends on torch.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libconsenrich_b200.so")
SOURCES = ["ssm_kernels.cu", "lean_kernels.cu", "apn_kernels.cu", "writer_kernels.cu", "background_kernels.cu", "munc_kernels.cu", "cabi.cu"]
HEADERS = ["ssm_math.cuh", "ssm_kernels.cuh", "lean_kernels.cuh", "writer_kernels.cuh", "background_kernels.cuh", "munc_kernels.cuh", os.path.join(ROOT, "include", "consenrich_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libconsenrich_b200.so cannot be built (there is no CPU build)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("CB200_EXTRA_NVCC", "").split()  # e.g. -DSCAN_MIN_CTAS=3 for experiments
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, *[os.path.join(CSRC, s) for s in SOURCES], "-o", LIB_PATH]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

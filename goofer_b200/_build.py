"""Build recipe of libgoofer_b200.so (nvcc, sm_100a only).  Called by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
# GOOFER_B200_LIB: load another build of the same sources (tools/build_variants.py: tuning experiments on the GPU box)
LIB_PATH = os.environ.get("GOOFER_B200_LIB") or os.path.join(LIB_DIR, "libgoofer_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return [os.path.join(CSRC, "goofer_b200.cu"), os.path.join(CSRC, "plan.cpp")]


def _newest_source_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for dp, _, fs in os.walk(root):
            for f in fs:
                m = max(m, os.path.getmtime(os.path.join(dp, f)))
    return m


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree.  Needs nvcc (cross-compiles without a GPU)."""
    if os.environ.get("GOOFER_B200_LIB"):
        return LIB_PATH
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libgoofer_b200.so cannot be built (there is no CPU fallback)")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    subprocess.check_call(cmd)
    return LIB_PATH

"""Build libsp_nnue.so (the C-ABI library, include/sp_nnue.h) in-tree with nvcc for sm_100a.

    python -m stormphrax_b200.build [--force]

The library is self-contained (static cudart, no torch): kernels + C-ABI + host utilities.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libsp_nnue.so")

CUDA_SOURCES = ["kernels.cu", "capi.cu", "selfplay_gpu.cu"]
HOST_SOURCES = ["host/position.cpp", "host/host_capi.cpp", "host/nnue_state.cpp", "host/selfplay.cpp", "host/net_loader.cpp"]
HEADERS = ["kernels.cuh", "sp_features.h", "sp_delta.h", "host/position.h", "host/nnue_state.h", "host/selfplay.h", "host/rng.h", "host/net_loader.h", "capi_slots.inc",
           "../../include/sp_nnue.h", "../../include/sp_types.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unknown-pragmas",
    "--expt-relaxed-constexpr", "--extended-lambda",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libsp_nnue.so cannot be built")
    return exe


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in CUDA_SOURCES + HOST_SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines: dict | None = None, out: str | None = None) -> str:
    """`defines` / `out` build a tuning variant (see the SP_* knobs at the top of kernels.cu)."""
    target = out or LIB_PATH
    if not force and not defines and not out and not _stale():
        return LIB_PATH
    os.makedirs(os.path.dirname(target), exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in CUDA_SOURCES + HOST_SOURCES if os.path.exists(os.path.join(CSRC, f))]
    cmd = [nvcc(), *NVCC_FLAGS, *[f"-D{k}={v}" for k, v in (defines or {}).items()], "-shared", "-o", target, *srcs, "-lpthread", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

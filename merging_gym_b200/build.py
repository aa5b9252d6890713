"""Builds `libmerging_b200.so` in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m merging_gym_b200.build [--force] [--verbose]

The shared library is a plain C-ABI artefact (include/merging_b200.h); it is git-ignored
but travels with the working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

from ._paths import CSRC, INCLUDE_DIR, LIB_DIR, LIB_PATH

SOURCES = [os.path.join(CSRC, "merge_kernels.cu"), os.path.join(CSRC, "mlp_kernels.cu"), os.path.join(CSRC, "mlp_tc_kernels.cu"),
           os.path.join(CSRC, "mlp_tc16_kernels.cu"),
           os.path.join(CSRC, "record_kernels.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "merge_device.cuh"), os.path.join(CSRC, "policy_env.cuh"), os.path.join(CSRC, "tc_common.cuh"),
                  os.path.join(CSRC, "abi_common.h"),
                  os.path.join(INCLUDE_DIR, "merging_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

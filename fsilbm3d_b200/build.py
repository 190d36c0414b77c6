"""Builds libfsilbm_b200.so in-tree with nvcc for sm_100a (B200).  No JIT cache: the .so sits next
to the sources so it travels with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfsilbm_b200.so")
SOURCES = ["fluid_kernels.cu", "ibm_kernels.cu", "refine_kernels.cu", "io_kernels.cu", "fsilbm_api.cu"]
HEADERS = ["d3q19.cuh", "kernels.h", os.path.join("..", "..", "include", "fsilbm.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--threads", "0",
    # no fused multiply-add: the reference's x86-64 gfortran build has none, and the parity tests
    # compare against an oracle compiled with -ffp-contract=off (see DESIGN.md "Arithmetic")
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-ccbin", "/usr/bin/g++", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libfsilbm_b200.so")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))

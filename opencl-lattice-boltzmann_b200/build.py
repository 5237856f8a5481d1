"""Builds everything native, in-tree: liblbm_b200.so (nvcc, sm_100a), the C host
``d2q9-bgk`` (gcc) and the oracle's C restatement (test infrastructure)."""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC,-Wall"]


def build_all(verbose: bool = False) -> None:
    """`make all` at the repo root (the Makefile holds the exact nvcc/gcc command lines:
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a ... -shared)."""
    env = dict(os.environ)
    proc = subprocess.run(["make", "-C", ROOT, "--no-print-directory", "all"], env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("native build failed (see output above)")

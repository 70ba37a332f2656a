"""Build libb200yolo.so in-tree with nvcc for sm_100a (no torch headers involved:
the library is a plain C-ABI shared object loaded with ctypes)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200yolo.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # one fp32 rounding per reference op; FMA only where written
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    hdr = os.path.join(os.path.dirname(HERE), "include", "b200yolo.h")
    return any(os.path.getmtime(f) > t for f in sources() + [hdr])


def find_nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.access(cand, os.X_OK):
            return cand
    raise RuntimeError("nvcc not found: libb200yolo.so cannot be built")


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", LIB, os.path.join(CSRC, "b200yolo.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))

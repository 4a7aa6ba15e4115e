"""Build libvermeer_b200.so (CUDA kernels + C ABI + host mirror) in-tree with nvcc for sm_100a.

    python -m vermeer_b200.build [--force]

nvcc cross-compiles without a GPU. Flags that matter for parity: -fmad=false (Go/amd64 never contracts
a*b+c; the traversal arithmetic must round like the reference's SSE code) and -ffp-contract=off for the host
mirror of PreRender. -lineinfo keeps ncu's source page usable.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libvermeer_b200.so")

SOURCES = [
    "context.cu", "kernels_trace.cu", "render.cu", "texture.cu", "build_bvh.cu",
    "host/builder.cpp", "host/nodes.cpp", "host/vh_capi.cpp", "host/vnf.cpp",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-pthread,-Wall,-Wno-unused-function",
    "-shared",
]


def _deps():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    out.append(os.path.join(HERE, "..", "include", "vermeer_gpu.h"))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in _deps()):
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("VG_EXTRA_NVCC_FLAGS", "").split()   # tuning experiments only (e.g. -DVG_REFILL_BELOW=20)
    out = os.environ.get("VG_SO_OUT", SO)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libvermeer_b200.so")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build libvermeer_b200.so (CUDA kernels + C ABI + host mirror) in-tree with nvcc for sm_100a.

    python -m vermeer_b200.build [--force] [-v]

nvcc cross-compiles without a GPU. Flags that matter for parity: -fmad=false (Go/amd64 never contracts
a*b+c; the traversal arithmetic must round like the reference's SSE code) and -ffp-contract=off for the host
mirror of PreRender. -lineinfo keeps ncu's source page usable.

Every source is compiled to its own object (in parallel, cached by modification time under build/) and the objects are
linked into the shared library; NCCL is NOT linked: csrc/comm.cu binds libnccl.so.2 with dlopen when vg_comm_init is
first called, so the library loads on hosts without NCCL.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libvermeer_b200.so")
OBJDIR = os.path.join(HERE, "build")

SOURCES = [
    "render.cu", "kernels_trace.cu", "context.cu", "texture.cu", "build_bvh.cu", "comm.cu", "peaks.cu",
    "host/builder.cpp", "host/nodes.cpp", "host/vh_capi.cpp", "host/vnf.cpp",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-pthread,-Wall,-Wno-unused-function",
]


def _deps():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    out.append(os.path.join(HERE, "..", "include", "vermeer_gpu.h"))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    out = os.environ.get("VG_SO_OUT", SO)
    extra = os.environ.get("VG_EXTRA_NVCC_FLAGS", "").split()   # tuning experiments only (e.g. -DVG_REFILL_BELOW=20)
    deps = _deps()
    newest = max(os.path.getmtime(d) for d in deps)
    if not force and not extra and os.path.exists(out) and newest <= os.path.getmtime(out):
        return out
    nvcc = os.environ.get("NVCC", "nvcc")
    tag = hashlib.sha1(" ".join(NVCC_FLAGS + extra).encode()).hexdigest()[:10]
    objdir = os.path.join(OBJDIR, tag)
    os.makedirs(objdir, exist_ok=True)
    # headers are shared by everything: an object is stale when ANY source or header is newer than it
    jobs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace("/", "_") + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            jobs.append((s, obj))

    def compile_one(job):
        s, obj = job
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, s)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write("== %s\n%s%s" % (s, r.stdout, r.stderr))
            failed |= r.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libvermeer_b200.so")
    objs = [os.path.join(objdir, s.replace("/", "_") + ".o") for s in SOURCES]
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC,-pthread", "-o", out] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libvermeer_b200.so")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

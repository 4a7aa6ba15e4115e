"""The C ABI's record layouts as a C compiler sees them (gcc on include/vermeer_gpu.h) against the Python mirrors the tests and the
bench pass through ctypes: a drift between the header and vermeer_b200/host.py would silently corrupt rays, hits or statistics."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_sizes(tmp_path, names, members):
    src = tmp_path / "abi.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vermeer_gpu.h"', "int main(void) {"]
    for n in names:
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (n, n))
    for s, m in members:
        lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (s, m, s, m))
    lines += ["  return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)]).decode().split()
    return {out[i]: int(out[i + 1]) for i in range(0, len(out), 2)}


def test_header_is_plain_c_and_matches_the_python_mirrors(tmp_path):
    sys.path.insert(0, ROOT)
    from vermeer_b200 import host
    names = ["VgRay", "VgRayPD", "VgHit", "VgHitCompact", "VgNode", "VgMotionNode", "VgStats", "VgCamera", "VgMaterial", "VgPeaks"]
    members = [("VgStats", "render_ms"), ("VgStats", "max_stack_depth"), ("VgStats", "shadow_level0_kernel"), ("VgHit", "prim"),
               ("VgHitCompact", "slot"), ("VgRay", "tmax"), ("VgRayPD", "d")]
    c = _c_sizes(tmp_path, names, members)
    assert c["VgRay"] == host.RAY_DTYPE.itemsize == 32
    assert c["VgRayPD"] == host.RAYPD_DTYPE.itemsize == 24          # three 8-byte loads per record on the device
    assert c["VgHit"] == host.HIT_DTYPE.itemsize == 32
    assert c["VgHitCompact"] == host.HITC_DTYPE.itemsize == 16
    assert c["VgNode"] == host.NODE_DTYPE.itemsize == 128            # qbvh.Node: one 128-byte line
    assert c["VgMotionNode"] == host.MNODE_DTYPE.itemsize
    assert c["VgStats"] == C.sizeof(host.VgStats)
    assert c["VgCamera"] == C.sizeof(host.VgCamera)
    assert c["VgMaterial"] == C.sizeof(host.VgMaterial)
    assert c["VgPeaks"] == C.sizeof(host.VgPeaks)
    for s, m in members:
        if s == "VgStats":
            assert c["%s.%s" % (s, m)] == getattr(host.VgStats, m).offset, m
    assert c["VgHit.prim"] == host.HIT_DTYPE.fields["prim"][1]
    assert c["VgHitCompact.slot"] == host.HITC_DTYPE.fields["slot"][1]
    assert c["VgRay.tmax"] == host.RAY_DTYPE.fields["tmax"][1]
    assert c["VgRayPD.d"] == host.RAYPD_DTYPE.fields["d"][1]


def _build_c_host(tmp_path, built_library):
    exe = tmp_path / "c_host"
    libdir = os.path.join(ROOT, "vermeer_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                           os.path.join(ROOT, "examples", "c_host.c"), "-L", libdir, "-lvermeer_b200", "-Wl,-rpath," + libdir, "-lm"])
    return exe


def test_plain_c_caller_of_the_host_layer(tmp_path, built_library):
    """examples/c_host.c, a C99 program that includes only include/vermeer_gpu.h and links the shared library: parses a .vnf
    scene, runs PreRender and reads the scene tree back (no GPU needed; the device half reports that there is none)."""
    exe = _build_c_host(tmp_path, built_library)
    r = subprocess.run([str(exe), os.path.join(ROOT, "examples", "cornell64.vnf")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host: 64x64, MaxIter 16, 9 geoms" in r.stdout, r.stdout


import pytest  # noqa: E402


@pytest.mark.gpu
def test_plain_c_caller_renders_on_the_device(tmp_path, built_library):
    exe = _build_c_host(tmp_path, built_library)
    r = subprocess.run([str(exe), os.path.join(ROOT, "examples", "cornell64.vnf")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "device: 4 iterations" in r.stdout, r.stdout
    rays = int(r.stdout.split("device: 4 iterations, ")[1].split(" rays")[0])
    assert rays >= 4 * 64 * 64                       # at least one TraceProbe per pixel and iteration
    mean = float(r.stdout.strip().split("mean pixel ")[1])
    assert 0.01 < mean < 10.0

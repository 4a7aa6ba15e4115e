"""The oracle against the REAL reference, for the machine that has a Go toolchain (oracle/go_ref/README.md is the recipe).
Skipped until oracle/_ref/go/<scene>/frame.float exist: this image cannot build Go, which is why DESIGN.md says "parity unpinned".
The input half of the recipe (scene files + ray batches) is exercised here unconditionally."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "oracle", "_ref", "go")
NAMES = ["cornell", "heightfield", "motion", "spheres"]


def test_inputs_of_the_recipe_are_reference_valid(built_library, tmp_path):
    """make_inputs.py writes scenes the .vnf reader (reference syntax) accepts without a message, and 32-byte ray records."""
    from vermeer_b200.host import HostScene, RAY_DTYPE
    sys.path.insert(0, os.path.join(ROOT, "oracle", "go_ref"))
    import make_inputs
    sc = make_inputs.SCENES["cornell"]()
    from vermeer_b200 import scenes
    h = HostScene.from_vnf(text=scenes.to_vnf(sc))
    assert h.parse_errors == 0
    assert RAY_DTYPE.itemsize == 32


@pytest.mark.parametrize("name", NAMES)
def test_oracle_equals_go_reference(oracle_lib, name):
    d = os.path.join(GO, name)
    if not os.path.exists(os.path.join(d, "frame.float")):
        pytest.skip("no Go run of the reference here (oracle/go_ref/README.md)")
    sys.path.insert(0, os.path.join(ROOT, "oracle", "go_ref"))
    import make_inputs
    from oracle.binding import Oracle
    from vermeer_b200.host import RAY_DTYPE
    sc = make_inputs.SCENES[name]()
    ora = Oracle(sc, motion_ref_compat=True)                       # the reference as it is, quirk (b) included
    tab = np.fromfile(os.path.join(d, "scramble.bin"), np.uint64).reshape(-1, 6)
    ora.set_scramble(tab)
    secs, rays, shadow = open(os.path.join(d, "stats.txt")).read().split()
    fo, so = ora.render(0, 4, nthreads=4)
    fg = np.fromfile(os.path.join(d, "frame.float"), np.float32).reshape(fo.shape)
    ok = np.isfinite(fo).all(-1) & np.isfinite(fg).all(-1)
    assert float(np.sqrt(((fo[ok] - fg[ok]) ** 2).mean())) <= 1e-6
    assert abs(so["rays"] - int(rays)) <= 1e-4 * int(rays)
    batch = np.fromfile(os.path.join(d, "rays.bin"), RAY_DTYPE)
    oh = ora.trace(batch)
    gh = np.fromfile(os.path.join(d, "hits.bin"), np.dtype([("t", "f4"), ("u", "f4"), ("v", "f4"), ("w", "f4"), ("prim", "i4"), ("geom", "i4"), ("nodesT", "i4"), ("leafsT", "i4")]))
    assert np.array_equal(oh["prim"] >= 0, gh["prim"] >= 0)
    hit = oh["prim"] >= 0
    assert np.array_equal(oh["prim"][hit], gh["prim"][hit])
    for f in ("t", "u", "v"):
        assert np.array_equal(oh[f][hit].view(np.uint32), gh[f][hit].view(np.uint32)), f
    assert np.array_equal(oh["nodesT"], gh["nodesT"])

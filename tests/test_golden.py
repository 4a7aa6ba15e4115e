"""Golden fixtures (tests/golden/, produced by tests/golden/make_golden.py from the oracle): the oracle must keep reproducing
them on CPU, and the CUDA path must reproduce them on the GPU."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import assert_hits_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _scene():
    from vermeer_b200 import scenes
    return scenes.cornell_box(48, 48)


def test_qmc_golden(oracle_lib):
    L = oracle_lib
    k = json.load(open(os.path.join(GOLD, "qmc_kat.json")))
    for i, s, want in k["vdc_u"]:
        assert L.orc_vdc_u(i, s) == want
    for i, s, want in k["sobol_u"]:
        assert L.orc_sobol_u(i, s) == want
    rx, ry = C.c_double(), C.c_double()
    for f, px, py, idx, hx, hy in k["raster"]:
        assert L.orc_raster_xy(f, px, py, 0, 0, C.byref(rx), C.byref(ry)) == idx
        assert rx.value == float.fromhex(hx) and ry.value == float.fromhex(hy)


def test_oracle_reproduces_golden_hits():
    from oracle.binding import Oracle
    g = np.load(os.path.join(GOLD, "cornell48_v1.npz"))
    ora = Oracle(_scene())
    assert_hits_equal(ora.trace(g["rays"]), g["closest"], what="golden closest (oracle)")
    assert_hits_equal(ora.trace(g["shadow_rays"], any_hit=True), g["anyhit"], what="golden any-hit (oracle)")


@pytest.mark.gpu
def test_gpu_reproduces_golden_hits(built_library):
    from vermeer_b200.host import Device, HostScene
    g = np.load(os.path.join(GOLD, "cornell48_v1.npz"))
    dev = Device(0).upload(HostScene(_scene()).prerender())
    assert_hits_equal(dev.trace(g["rays"]), g["closest"], what="golden closest (gpu)")
    assert_hits_equal(dev.trace(g["shadow_rays"], any_hit=True), g["anyhit"], what="golden any-hit (gpu)")


@pytest.mark.gpu
def test_gpu_reproduces_golden_image(built_library):
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    g = np.load(os.path.join(GOLD, "cornell48_v1.npz"))
    sc = _scene()
    dev = Device(0).upload(HostScene(sc).prerender())
    dev.set_scramble(scenes.splitmix64_table(7, sc.XRes * sc.YRes))
    fg = dev.render(0, 4)
    fo = g["image"]
    ok = np.isfinite(fo).all(-1) & np.isfinite(fg).all(-1)
    assert np.sqrt(((fo[ok] - fg[ok]) ** 2).mean()) <= 1e-3
    st = dev.stats()
    assert abs(int(st["rays"]) - int(g["ray_count"][0])) <= 1e-3 * int(g["ray_count"][0])


def _texture_golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.texture_golden_inputs(), np.load(os.path.join(GOLD, "texture_v1.npz"))


def test_oracle_reproduces_golden_textures(oracle_lib):
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    (imgs, co), g = _texture_golden()
    assert np.array_equal(co, g["coords"])
    sc = scenes.cornell_box(8, 8, boxes=False)
    sc.textures = [scenes.Texture(k, v) for k, v in imgs.items()]
    ora = Oracle(sc)
    for k in imgs:
        for l, a in enumerate(ora.texture_levels(k)):
            assert np.array_equal(a, g["%s_level%d" % (k, l)])
        assert np.array_equal(ora.texture_sample(k, co, trilinear=False).view(np.uint32), g["%s_feline" % k].view(np.uint32))
        assert np.array_equal(ora.texture_sample(k, co, trilinear=True).view(np.uint32), g["%s_trilinear" % k].view(np.uint32))


@pytest.mark.gpu
def test_gpu_reproduces_golden_textures(built_library):
    from vermeer_b200.host import Device
    (imgs, co), g = _texture_golden()
    dev = Device(0)
    for k, img in imgs.items():
        tid = dev.texture_upload(img[::-1])
        for l, a in enumerate(dev.texture_levels(tid)):
            assert np.array_equal(a, g["%s_level%d" % (k, l)]), (k, l)
        for name, tri in (("feline", False), ("trilinear", True)):
            d = np.abs(dev.texture_sample(tid, co, trilinear=tri) - g["%s_%s" % (k, name)])
            assert d.max() <= 2e-5, (k, name, d.max())

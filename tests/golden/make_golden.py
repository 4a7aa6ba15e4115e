#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the oracle (the Go reference cannot run in this image, so these
vectors pin the oracle-as-built and the CUDA path against regressions; they are not reference outputs).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.binding import Oracle, lib  # noqa: E402
from vermeer_b200 import scenes  # noqa: E402
import ctypes as C  # noqa: E402


def golden_scene():
    return scenes.cornell_box(48, 48)


def main():
    L = lib()
    # 1. QMC known answers
    rx, ry = C.c_double(), C.c_double()
    qmc = {"vdc_u": [], "sobol_u": [], "raster": []}
    rng = np.random.default_rng(12345)
    for _ in range(64):
        i = int(rng.integers(0, 1 << 40))
        s = int(rng.integers(0, 1 << 63))
        qmc["vdc_u"].append([i, s, int(L.orc_vdc_u(i, s))])
        qmc["sobol_u"].append([i, s, int(L.orc_sobol_u(i, s))])
    for _ in range(64):
        f, px, py = int(rng.integers(1, 1 << 20)), int(rng.integers(0, 4096)), int(rng.integers(0, 4096))
        idx = int(L.orc_raster_xy(f, px, py, 0, 0, C.byref(rx), C.byref(ry)))
        qmc["raster"].append([f, px, py, idx, rx.value.hex(), ry.value.hex()])
    json.dump(qmc, open(os.path.join(HERE, "qmc_kat.json"), "w"), indent=0)

    # 2. a ray batch through the Cornell box: camera rays of iteration 3 + random interior rays, closest and any-hit
    sc = golden_scene()
    ora = Oracle(sc)
    ora.set_scramble(scenes.splitmix64_table(7, sc.XRes * sc.YRes))
    cam = ora.camera_rays(3)
    r = np.zeros(2000, cam.dtype)
    r["o"] = rng.uniform((-0.95, 0.05, -0.95), (0.95, 1.9, 0.95), (2000, 3)).astype(np.float32)
    d = rng.normal(size=(2000, 3))
    r["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    r["tmax"] = np.inf
    rays = np.concatenate([cam, r])
    closest = ora.trace(rays)
    sh = rays.copy()
    sh["tmax"] = 1.25
    anyhit = ora.trace(sh, any_hit=True)
    fb, st = ora.render(0, 4, nthreads=1)
    np.savez_compressed(os.path.join(HERE, "cornell48_v1.npz"), rays=rays, closest=closest, shadow_rays=sh, anyhit=anyhit,
                        image=fb, ray_count=np.asarray([st["rays"], st["shadow_rays"]]))
    print("wrote qmc_kat.json, cornell48_v1.npz (%d rays)" % len(rays))
    make_texture_golden()


def texture_golden_inputs():
    """Seeded inputs of texture_v1.npz: two images (one odd-sized) and a batch of lookups."""
    rng = np.random.default_rng(77)
    imgs = {"sq": rng.integers(0, 256, (32, 32, 3), dtype=np.uint8), "odd": rng.integers(0, 256, (13, 21, 3), dtype=np.uint8)}
    n = 512
    co = np.zeros((n, 8), np.float32)
    co[:, 0:2] = rng.uniform(-1, 2, (n, 2))
    r = 10.0 ** rng.uniform(-3, 0, n)
    ratio = 10.0 ** rng.uniform(0, 1.5, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    co[:, 2] = np.cos(ang) * r
    co[:, 3] = np.sin(ang) * r
    co[:, 4] = -np.sin(ang) * r / ratio
    co[:, 5] = np.cos(ang) * r / ratio
    co[:, 6:8] = 1.0
    return imgs, co


def make_texture_golden():
    """3. the texture subsystem: mip pyramids (bytes) and both filters on a seeded batch of lookups."""
    from vermeer_b200 import scenes
    imgs, co = texture_golden_inputs()
    sc = scenes.cornell_box(8, 8, boxes=False)
    sc.textures = [scenes.Texture(k, v) for k, v in imgs.items()]
    ora = Oracle(sc)
    out = {"coords": co}
    for k in imgs:
        for l, a in enumerate(ora.texture_levels(k)):
            out["%s_level%d" % (k, l)] = a
        out["%s_feline" % k] = ora.texture_sample(k, co, trilinear=False)
        out["%s_trilinear" % k] = ora.texture_sample(k, co, trilinear=True)
    np.savez_compressed(os.path.join(HERE, "texture_v1.npz"), **out)
    print("wrote texture_v1.npz")


if __name__ == "__main__":
    main()

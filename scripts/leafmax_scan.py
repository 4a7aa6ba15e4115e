"""Opt-in non-parity mode: the C2 frame and the incoherent closest-hit batch with the reference's builder run at leafMax 16 (parity), 8, 4, 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene
scene = bench.build_scene("c2")
tab = scenes.splitmix64_table(1, 1920 * 1080)
ref = None
for lm in (16, 8, 4, 2):
    host = HostScene(scene, leaf_max=lm).prerender()
    dev = Device(0).upload(host)
    dev.set_scramble(tab)
    dev.set_option("iters_per_batch", 32)
    for _ in range(2):
        dev.clear(); dev.render(0, 64, fetch=False)
    st = dev.stats()
    inc = bench.incoherent_batch(scene, host.camera(), lambda r: dev.trace(r))
    tb = bench.time_batch(dev, torch, inc, 5, 3, compact_e2e=False)
    _, aidx = host.mesh_idxp(0)
    h = tb["hits"]
    face = np.where(h["prim"] >= 0, aidx[np.maximum(h["prim"], 0)], -1)
    face = np.where(h["geom"] == 0, face, -2 - h["geom"])
    if ref is None:
        ref = (h["t"].copy(), face.copy())
    same_t = float((h["t"].view(np.uint32) == ref[0].view(np.uint32)).mean()); same_f = float((face == ref[1]).mean())
    print("leaf_max %2d: nodes %6d frame %.2f ms (closest %.2f shadow %.2f shade %.2f) | incoherent %.1f Mrays/s nodesT %.1f trisT %.1f | same t %.6f same face %.6f" % (
        lm, host.mesh_info(0)["nodes"], st["render_ms"], st["closest_ms"], st["shadow_ms"], st["shade_ms"], tb["rays"] / tb["ms_per_step"] / 1e3, tb["nodesT_per_ray"], tb["trisT_per_ray"], same_t, same_f))

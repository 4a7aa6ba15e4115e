"""Where the end-to-end step (bench.py's e2e leg, N=1) spends its time: vg_set_scramble (host gather + H2D), vg_clear_framebuffer,
vg_render without and with the framebuffer D2H."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene
sc = scenes.heightfield_scene(1920, 1080, nq=708)
dev = Device(0).upload(HostScene(sc).prerender())
tab = scenes.splitmix64_table(1, 1920 * 1080)
dev.set_scramble(tab)
dev.set_option("iters_per_batch", 8)
for k, v in [kv.split("=") for kv in os.environ.get("VG_OPTIONS", "").split(",") if kv]:
    dev.set_option(k, int(v))
dev.render(0, 8, fetch=False)
def t(f, n=5):
    f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("set_scramble %.2f ms" % t(lambda: dev.set_scramble(tab)))
print("clear        %.2f ms" % t(dev.clear))
print("render(64) no fetch %.2f ms (device %.2f)" % (t(lambda: dev.render(0, 64, fetch=False), 3), dev.stats()["render_ms"]))
print("render(64) fetch    %.2f ms" % t(lambda: dev.render(0, 64, fetch=True), 3))
st = dev.stats()
print({k: st[k] for k in ("render_ms", "closest_ms", "shadow_ms")})

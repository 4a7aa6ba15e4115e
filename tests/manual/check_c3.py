"""C3-scale sanity: 1024 displaced spheres x 9800 tris (10.04 M triangles) under the scene-level QBVH, mirror chains.
Checks host PreRender time, device memory, traversal parity against the oracle on a sample of camera rays, and throughput."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene

W, H = 1920, 1080
t = time.time(); sc = scenes.sphere_field_scene(W, H); print("generate %.1fs, %d tris, %d meshes" % (time.time() - t, sc.num_tris, len(sc.meshes)))
t = time.time(); host = HostScene(sc).prerender(); print("host prerender %.2fs" % (time.time() - t))
t = time.time(); dev = Device(0).upload(host); print("upload %.2fs" % (time.time() - t))
tab = scenes.splitmix64_table(3, W * H)
dev.set_scramble(tab)
dev.set_option("iters_per_batch", 4)
dev.render(0, 4, fetch=False)
dev.reset_stats(); dev.clear()
t = time.time(); fb = dev.render(0, 16); wall = time.time() - t
st = dev.stats()
print("GPU 16 spp: %.1f ms device, %.2f s wall, %d rays (%d shadow) -> %.1f Mrays/s; closest %.1f ms / shadow %.1f ms; nodesT/ray %.1f trisT/ray %.1f" % (
    st["render_ms"], wall, st["rays"], st["shadow_rays"], st["rays"] / st["render_ms"] / 1e3, st["closest_ms"], st["shadow_ms"],
    st["nodes_t"] / max(1, st["rays"] - st["shadow_rays"]), st["tris_t"] / max(1, st["rays"] - st["shadow_rays"])))
print("image mean %.5f, nan px %d" % (np.nanmean(fb), (~np.isfinite(fb).all(-1)).sum()))
if "--oracle" in sys.argv:
    from oracle.binding import Oracle
    t = time.time(); ora = Oracle(sc); print("oracle prerender %.1fs" % (time.time() - t))
    ora.set_scramble(tab)
    rays = ora.camera_rays(1, 0, 500, W, 32)
    g = dev.trace(rays); o = ora.trace(rays, nthreads=8)
    same = all(np.array_equal(g[f], o[f]) for f in ("prim", "geom", "nodesT", "trisT")) and np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
    print("traversal parity on %d camera rays: %s (hit %.3f)" % (len(rays), same, (g["prim"] >= 0).mean()))
    t = time.time(); fo, so = ora.render(0, 1, nthreads=os.cpu_count()); print("oracle 1 spp: %.1fs, %d rays -> %.2f Mrays/s on %d threads" % (so["seconds"], so["rays"], so["rays"] / so["seconds"] / 1e6, os.cpu_count()))

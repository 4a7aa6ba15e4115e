"""GPU vs oracle image comparison on small configs (debug helper; the pytest versions live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle.binding import Oracle
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene

def compare(name, sc, iters, seed=1, nthreads=8, **opts):
    tab = scenes.splitmix64_table(seed, sc.XRes * sc.YRes)
    ora = Oracle(sc, motion_ref_compat=False); ora.set_scramble(tab)
    t = time.time(); fo, so = ora.render(0, iters, nthreads=nthreads); to = time.time() - t
    host = HostScene(sc).prerender()
    dev = Device(0).upload(host); dev.set_scramble(tab)
    for k, v in opts.items(): dev.set_option(k, v)
    t = time.time(); fg = dev.render(0, iters); tg = time.time() - t
    st = dev.stats()
    both = np.isfinite(fo).all(-1) & np.isfinite(fg).all(-1)
    rmse = np.sqrt(((fo[both] - fg[both]) ** 2).mean())
    print("%-10s %dx%d it=%d  rmse=%.3e  max=%.3e  mean o/g=%.5f/%.5f  nan o/g=%d/%d  rays o/g=%d/%d shadow o/g=%d/%d  cpu %.2fs gpu %.3fs (%.1f ms dev)" % (
        name, sc.XRes, sc.YRes, iters, rmse, np.abs(fo[both] - fg[both]).max(), fo[both].mean(), fg[both].mean(),
        (~np.isfinite(fo).all(-1)).sum(), (~np.isfinite(fg).all(-1)).sum(), so["rays"], st["rays"], so["shadow_rays"], st["shadow_rays"], to, tg, st["render_ms"]))
    return fo, fg

compare("cornell", scenes.cornell_box(64, 64), 4)
compare("cornell", scenes.cornell_box(128, 128), 16, iters_per_batch=5)
compare("heightf", scenes.heightfield_scene(160, 90, nq=100), 8)
compare("spheres", scenes.sphere_field_scene(96, 96, nmesh=16, slices=16, stacks=17), 4)
compare("sph-nolast", scenes.sphere_field_scene(96, 96, nmesh=16, slices=16, stacks=17), 4, trace_last_level=0)
compare("motion", scenes.heightfield_scene(96, 96, nq=60, motion=True), 4)

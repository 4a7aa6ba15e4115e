"""FAST (algebraic, single precision) vs precise (float64 trig, literal transcription) shading against the oracle:
RMSE, median |diff|, max |diff| and ray counts on the parity scenes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle.binding import Oracle
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene

def run(name, sc, iters):
    tab = scenes.splitmix64_table(1, sc.XRes * sc.YRes)
    ora = Oracle(sc, motion_ref_compat=False); ora.set_scramble(tab)
    fo, so = ora.render(0, iters, nthreads=os.cpu_count())
    for precise in (1, 0):
        dev = Device(0).upload(HostScene(sc).prerender()); dev.set_scramble(tab); dev.set_option("precise_trig", precise)
        fg = dev.render(0, iters); st = dev.stats()
        ok = np.isfinite(fo).all(-1) & np.isfinite(fg).all(-1)
        d = np.abs(fo[ok] - fg[ok])
        print("%-12s precise=%d rmse %.3e median %.2e p99 %.2e max %.2e  nan(o/g) %d/%d rays o/g %d/%d shadow %d/%d  mean %.4f" % (
            name, precise, np.sqrt((d ** 2).mean()), np.median(d), np.quantile(d, 0.99), d.max(), (~np.isfinite(fo).all(-1)).sum(), (~np.isfinite(fg).all(-1)).sum(),
            so["rays"], st["rays"], so["shadow_rays"], st["shadow_rays"], fo[ok].mean()))

run("heightfield", scenes.heightfield_scene(192, 108, nq=120), 8)
run("cornell", scenes.cornell_box(128, 128), 16)
run("motion", scenes.heightfield_scene(128, 96, nq=60, motion=True), 8)
run("spheres", scenes.sphere_field_scene(96, 96, nmesh=16, slices=16, stacks=17), 32)

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle.binding import Oracle
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene

def run(sc, iters, tag):
    tab = scenes.splitmix64_table(1, sc.XRes * sc.YRes)
    ora = Oracle(sc); ora.set_scramble(tab)
    fo, so = ora.render(0, iters, nthreads=8)
    dev = Device(0).upload(HostScene(sc).prerender()); dev.set_scramble(tab)
    fg = dev.render(0, iters)
    st = dev.stats()
    ok = np.isfinite(fo).all(-1) & np.isfinite(fg).all(-1)
    d = np.abs(fo - fg).max(-1); d[~ok] = 0
    print(tag, "rmse %.3e  median %.2e  frac>1e-3 %.4f  frac>1e-2 %.4f  max %.3f  rays o/g %d/%d  mean o/g %.5f/%.5f" % (
        np.sqrt(((fo[ok]-fg[ok])**2).mean()), np.median(d), (d>1e-3).mean(), (d>1e-2).mean(), d.max(), so["rays"], st["rays"], fo[ok].mean(), fg[ok].mean()))
    ys, xs = np.nonzero(d > 1e-2)
    print("   rows of bad px:", np.bincount(ys // 16, minlength=8), " cols:", np.bincount(xs // 16, minlength=8))
    return fo, fg

mir = scenes.ShaderStd("mirror", DiffuseColour=(0.5, 0.5, 0.5), DiffuseStrength=0.3, Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.7, Spec1Roughness=0.0)
sc = scenes.cornell_box(128, 128); sc.shaders.append(mir); sc.meshes[0].Shader = ["mirror"]
run(sc, 1, "floor mirror, 1 it ")
run(sc, 16, "floor mirror, 16 it")
sc = scenes.cornell_box(128, 128, boxes=False); sc.shaders.append(mir); sc.meshes[0].Shader = ["mirror"]
run(sc, 4, "floor mirror no boxes")
sc = scenes.heightfield_scene(128, 128, nq=8); sc.shaders[0] = scenes.ShaderStd("ground", DiffuseColour=(0.5, 0.5, 0.5), DiffuseStrength=0.3, Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.7, Spec1Roughness=0.0)
run(sc, 4, "mirror heightfield")
sc = scenes.heightfield_scene(128, 128, nq=8); sc.shaders[0] = scenes.ShaderStd("ground", DiffuseColour=(0.5, 0.5, 0.5), DiffuseStrength=0.3, Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.7, Spec1Roughness=0.0)
sc.lights = sc.lights[:1]
run(sc, 4, "mirror heightfield, ONE light")
sc = scenes.cornell_box(128, 128); sc.shaders.append(mir); sc.meshes[0].Shader = ["mirror"]; sc.meshes[2].Shader = ["mirror"]; sc.lights = sc.lights[:1]
run(sc, 16, "floor+back mirror, ONE light")
sc = scenes.sphere_field_scene(96, 96, nmesh=16, slices=16, stacks=17); sc.lights = sc.lights[:1]
run(sc, 64, "sphere field, ONE light")

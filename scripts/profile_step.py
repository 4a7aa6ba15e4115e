"""One 1080p C2 render of 8 iterations (2 batches) — short target for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene
motion = "--motion" in sys.argv
sc = scenes.sphere_field_scene(1920, 1080) if "--c3" in sys.argv else scenes.heightfield_scene(1920, 1080, nq=708, motion=motion)
if "--tex" in sys.argv:   # the C2T bench scene: a 1024x1024 Feline map on the ground
    import numpy as np
    m = sc.meshes[0]
    xz = m.Verts[0][:, [0, 2]]
    m.UV = ((xz - xz.min(0)) / (xz.max(0) - xz.min(0)) * 8.0).astype(np.float32)
    sc.textures = [scenes.Texture("ground.png", scenes._test_texture(1024, 1024, 21))]
    [s for s in sc.shaders if s.Name == m.Shader[0]][0].DiffuseColour = "ground.png"
host = HostScene(sc).prerender()
dev = Device(0).upload(host)
dev.set_scramble(scenes.splitmix64_table(1, 1920 * 1080))
iters = 8
for a in sys.argv:
    if a.startswith("--iters="):
        iters = int(a.split("=")[1])
    if a.startswith("--opt="):   # --opt=name=value
        _, k, v = a.split("=")
        dev.set_option(k, int(v))
dev.render(0, iters, fetch=False)
st = dev.stats()
print(st)
for a in sys.argv:
    if a.startswith("--stats-out="):
        import json
        json.dump(dict(st, iters=iters), open(a.split("=", 1)[1], "w"))

"""Scene preparation times (host PreRender, vh_upload = flatten + H2D) for a bench config: python scripts/prep_time.py c2|c3"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
t = time.time()
sc = scenes.sphere_field_scene(1920, 1080) if which == "c3" else scenes.heightfield_scene(1920, 1080, nq=708)
print("%s: generate %.2f s (%d triangles)" % (which, time.time() - t, sc.num_tris), flush=True)
dev = Device(0)
for rep in range(2):
    t = time.time(); h = HostScene(sc); t_nodes = time.time() - t
    t = time.time(); h.prerender(); t_pre = time.time() - t
    t = time.time(); dev.upload(h); t_up = time.time() - t
    print("%s: nodes %.3f s, prerender %.3f s, upload %.3f s" % (which, t_nodes, t_pre, t_up), flush=True)

"""Host vs device QBVH build time for the bench scenes' meshes (vh_prerender vs vh_prerender_device), and the traversal rate on each tree."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
sc = scenes.sphere_field_scene(1920, 1080) if which == "c3" else scenes.heightfield_scene(1920, 1080, nq=708)
dev = Device(0)
HostScene(scenes.heightfield_scene(64, 64, nq=80)).prerender(device=dev)   # warm up (context, module load)
for name, d in (("host", None), ("device", dev), ("host", None), ("device", dev)):
    h = HostScene(sc)
    t = time.time()
    h.prerender(device=d)
    dt = time.time() - t
    print("%s %s prerender %.1f ms (%d triangles)" % (which, name, dt * 1e3, sc.num_tris), flush=True)
st0 = dev.stats()["kernel_launches"]
h = HostScene(sc)
h.prerender(device=dev)
print("device build kernel launches:", dev.stats()["kernel_launches"] - st0)

# the raw entry point on the same mesh: boxes/centroids prepared by the caller, as qbvh.BuildAccel takes them
m = sc.meshes[0] if which != "c3" else sc.meshes[1]
v = m.Verts[0]
f = (m.FaceIdx if m.FaceIdx is not None else np.arange(len(v))).reshape(-1, 3)
if m.PolyCount is not None:
    f = None
if f is not None:
    p = v[f]
    boxes = np.concatenate([p.min(1), p.max(1)], 1).astype(np.float32)
    cent = ((p[:, 0] + p[:, 1] + p[:, 2]) / np.float32(3)).astype(np.float32)
    for rep in range(3):
        t = time.time()
        nodes, idx, b6 = dev.build_qbvh(boxes, cent)
        print("vg_build_qbvh(%d prims): %.1f ms wall, %d nodes" % (len(boxes), (time.time() - t) * 1e3, len(nodes)), flush=True)

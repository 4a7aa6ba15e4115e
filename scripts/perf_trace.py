"""Traversal-kernel throughput on the C2 scene (1M-triangle heightfield): primary, incoherent and shadow batches,
rays resident in HBM (vg_trace_batch_device). Prints Mrays/s, mean NodesT/TrisT and the algorithmic GB/s."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene, RAY_DTYPE, HIT_DTYPE

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 708
W, H = 1920, 1080
sc = scenes.heightfield_scene(W, H, nq=nq)
t = time.time(); host = HostScene(sc).prerender(); print("prerender %.2fs" % (time.time() - t))
dev = Device(0).upload(host)
if os.environ.get("VG_TRAVERSAL"): dev.set_option("traversal", int(os.environ["VG_TRAVERSAL"]))
cam_m, ttf, asp = host.camera()

# primary rays: pinhole through pixel centres (host numpy; not the QMC sampler)
ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
sx = (-1 + 2 * (xs + 0.5) / W).astype(np.float32); sy = -(-1 + 2 * (ys + 0.5) / H).astype(np.float32)
M = cam_m.reshape(4, 4).T
d = np.stack([sx * ttf, sy * (ttf / asp), -np.full_like(sx, sc.camera.Focal)], -1).reshape(-1, 3)
dw = d @ M[:3, :3].T
dw /= np.linalg.norm(dw, axis=1, keepdims=True)
rays = np.zeros(W * H, RAY_DTYPE)
rays["o"] = M[:3, 3]; rays["d"] = dw.astype(np.float32); rays["tmax"] = np.inf

def run(name, rays, any_hit=False, reps=5):
    n = len(rays)
    d_r = torch.from_numpy(rays.view(np.uint8).reshape(n, 32)).cuda()
    d_h = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    best = 1e9
    for i in range(reps + 2):
        dev.reset_stats()
        dev.trace_device(d_r.data_ptr(), n, d_h.data_ptr(), any_hit)
        st = dev.stats()
        if i >= 2: best = min(best, st["trace_ms"])
    hits = d_h.cpu().numpy().view(HIT_DTYPE).reshape(-1)
    nodes, tris = st["nodes_t"] / n, st["tris_t"] / n
    bytes_ray = 64 + 128 * nodes + 48 * tris
    print("%-12s n=%d  %.3f ms  %.1f Mrays/s  hit=%.3f  NodesT=%.2f TrisT=%.2f  B/ray=%.0f  alg %.1f GB/s" % (
        name, n, best, n / best / 1e3, (hits["prim"] >= 0).mean(), nodes, tris, bytes_ray, bytes_ray * n / best / 1e6))
    return hits

h = run("primary", rays)
inc = scenes.incoherent_rays(rays, h, seed=5)
h2 = run("incoherent", inc)
sh = inc.copy(); sh["tmax"] = 0.5
run("shadow(any)", sh, any_hit=True)
rng = np.random.default_rng(1); perm = rng.permutation(len(inc))
run("incoh-shuf", inc[perm])

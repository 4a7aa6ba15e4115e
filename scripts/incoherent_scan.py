"""Closest-hit rate of vg_trace_batch_device on the C2 incoherent batch as a function of the batch size (k x 1.48 M rays)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene, RAY_DTYPE
X, Y = 1920, 1080
sc = scenes.heightfield_scene(X, Y, nq=708)
host = HostScene(sc).prerender()
dev = Device(0).upload(host)
for k, v in [kv.split("=") for kv in os.environ.get("VG_OPTIONS", "").split(",") if kv]:
    dev.set_option(k, int(v))
cam_m, ttf, asp = host.camera()
M = cam_m.reshape(4, 4).T
def primary(jx, jy):
    ys, xs = np.meshgrid(np.arange(Y), np.arange(X), indexing="ij")
    sx = (-1 + 2 * (xs + jx) / X).astype(np.float32); sy = -(-1 + 2 * (ys + jy) / Y).astype(np.float32)
    d = np.stack([sx * ttf, sy * (ttf / asp), -np.full_like(sx, sc.camera.Focal)], -1).reshape(-1, 3) @ M[:3, :3].T
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros(X * Y, RAY_DTYPE); r["o"] = M[:3, 3]; r["d"] = d.astype(np.float32); r["tmax"] = np.inf
    return r
parts = []
for s, (jx, jy) in enumerate([(0.5, 0.5), (0.25, 0.75), (0.75, 0.25), (0.1, 0.4), (0.9, 0.6), (0.35, 0.15), (0.65, 0.85), (0.45, 0.05)]):
    p = primary(jx, jy)
    parts.append(scenes.incoherent_rays(p, dev.trace(p), seed=5 + s))
for k in (1, 2, 4, 8):
    inc = np.concatenate(parts[:k]); inc = inc[np.random.default_rng(1).permutation(len(inc))]
    n = len(inc)
    d_r = torch.from_numpy(inc.view(np.uint8).reshape(n, 32)).cuda(); d_h = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    best = 1e30
    for i in range(8):
        dev.reset_stats(); dev.trace_device(d_r.data_ptr(), n, d_h.data_ptr(), False)
        if i >= 3: best = min(best, dev.stats()["trace_ms"])
    print("k=%d rays=%d best %.3f ms -> %.0f Mrays/s" % (k, n, best, n / best / 1e3))

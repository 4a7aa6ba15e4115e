"""The bench's headline batch (incoherent closest-hit rays on the 1M-triangle scene), three device-resident passes — short target
for `ncu --set full -k regex:k_trace_batch`. Writes the ray count next to the capture (--stats-out=file.json)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from vermeer_b200.host import Device, HostScene
scene = bench.build_scene("c2")
host = HostScene(scene).prerender()
dev = Device(0).upload(host)
inc = bench.incoherent_batch(scene, host.camera(), lambda r: dev.trace(r))
n = len(inc)
d_r = torch.from_numpy(inc.view(np.uint8).reshape(n, 32)).cuda()
d_h = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
for _ in range(3):
    dev.trace_device(d_r.data_ptr(), n, d_h.data_ptr(), False)
st = dev.stats()
print(n, st["trace_ms"])
for a in sys.argv:
    if a.startswith("--stats-out="):
        json.dump({"rays_per_launch": n, "trace_ms": st["trace_ms"]}, open(a.split("=", 1)[1], "w"))

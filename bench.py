#!/usr/bin/env python3
"""bench.py — the benchmark of the accelerated path (contract: the task's bench section + SURVEY.md 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-cpu] [--configs c1,c2,c3,c4]

BASELINE.json's metric is "Mrays/s (closest-hit, incoherent) and samples/sec at 1/2/4/8 B200 vs host CPU", quoted on the
1M-triangle config. The ONE JSON line this prints therefore has

  value / metric   closest-hit traversal of an INCOHERENT ray batch on the 1M-triangle scene (configs[1]): cosine-hemisphere
                   bounce rays from the primary hit points, shuffled, identical rays on GPU and CPU. One STEP = one
                   TraceProbe pass (core/trace.go:26) over the whole batch, rays and hits resident in HBM
                   (vg_trace_batch_device); device time from CUDA events around the kernel. At N GPUs every rank traces a
                   batch of its own (independent rays, no collective): weak scaling.
  e2e              the same pass through vg_trace_batch with page-locked HOST buffers: H2D of the rays, traversal and D2H of
                   the hits inside the timed call (VG_TRACE_RAYS_PD 24-byte {P, D} rays up, VG_TRACE_COMPACT_HITS 16-byte hits
                   down; the 32-byte VgRay / VgHit figures are beside it, the hits are byte-identical in every mode).
  frame            the other half of the metric, samples/s: the whole hot path on the same scene (ray generation ->
                   closest-hit -> ShaderStd/light sampling -> any-hit -> accumulation), 1920x1080, 64 spp; Mrays/s by the
                   reference's definition (core/stats.go:21-24: every TraceProbe / render-loop time). At N GPUs the frame
                   is split by 32x32 tiles and gathered on rank 0 by the library (vg_gather_frame: NCCL): strong scaling.
                   e2e = vg_set_scramble (H2D) + vg_render + gather + D2H of the frame, every step.
  configs          the same frame measurement for every BASELINE.json config that fits the run: c1 (Cornell 512^2, 16 spp),
                   c2, c3 (10M triangles, mirror chains, 256 spp), c4 (MQBVH, 2 keys) and, at 8 GPUs, c5 (4K, 1024 spp),
                   each with its own cpu_baseline (short sample: it is a rate) and per-stage roofline.
  incoherent_wavefront  a second, harder incoherent batch: the level-1..3 mirror-bounce rays of the 10M-triangle scene taken
                   from the integrator's own ray queues (vg_captured_rays), shuffled; scene not L2-resident.
  incoherent_diffuse    BASELINE.json's third config by name: bounces 1-4 of cosine-hemisphere DIFFUSE paths on the same scene
                   (one camera pass; scenes.diffuse_bounce_rays), per bounce and together, CPU on the same rays, bit-identity.
  roofline         per traversal / shading stage: time per launch measured live (CUDA events inside vg_render), traffic per
                   ray from the committed ncu capture of this build (profiles/ncu_r02_metrics.json: dram, L2 and L1TEX bytes),
                   against HBM (MEASURED_PEAKS.json) and the L2 / L1 read peaks measured in this run (vg_measure_peaks);
                   plus the issue-slot x SIMD-lane utilisation of the capture, the ceiling the ncu data says binds.
  cpu_baseline     the oracle (C++ restatement of the reference path; the Go reference cannot be built here: no Go
                   toolchain) on the box's host cores, bounded samples; `per_core` and the reference-faithful variant
                   (min(10, cores) workers + two global atomic counters per ray: core/render.go:190, core/stats.go:26-33).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c1": dict(workload="C1: Cornell box (5 diffuse walls, 2 boxes, quad light = 2 TriLights), 512x512, 16 spp", spp=16, xres=512, yres=512),
    "c2": dict(workload="C2: 1M-triangle heightfield (1002528 tris), primary+shadow rays, 1920x1080, 64 spp", spp=64),
    "c3": dict(workload="C3: 10M-triangle displaced-sphere field (1024 meshes x 9800 tris), mirror chains, 1920x1080, 256 spp", spp=256),
    "c4": dict(workload="C4: motion-blurred 1M-triangle heightfield (MQBVH, 2 keys), 1920x1080, 64 spp", spp=64),
    "c2t": dict(workload="C2T: C2 with a 1024x1024 Feline texture map on the ground (UV = 8 tiles over the mesh), 1920x1080, 64 spp", spp=64),
    "c5": dict(workload="C5: 10M-triangle displaced-sphere field, mirror chains, 3840x2160, 1024 spp", spp=1024, xres=3840, yres=2160),
}
HEADLINE_WORKLOAD = ("C2-incoherent: closest-hit TraceProbe over cosine-hemisphere bounce rays from the primary hit points of the "
                     "1M-triangle heightfield (1002528 tris), 4 jittered 1920x1080 primary passes, shuffled")
METRIC = "Mrays/s (closest-hit, incoherent)"
NQ = 708
SCRAMBLE_SEED = 1
ITERS_PER_BATCH = 32   # wavefront batch depth at N=1 (a warp = one pixel x 32 iterations); scaled by N, capped at the frame's spp


def cfg_res(cfg):
    c = CONFIGS[cfg]
    return c.get("xres", 1920), c.get("yres", 1080), c["spp"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_metrics():
    p = os.path.join(ROOT, "profiles", "ncu_r02_metrics.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        # the samples under load are the ones that matter: an idle GPU between legs reports its idle clock
        hot = [s for s in sm if s >= 0.6 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(hot)) if hot else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
_SCENE_CACHE = {}


def build_scene(cfg):
    from vermeer_b200 import scenes
    xres, yres, _ = cfg_res(cfg)
    key = ("c3" if cfg in ("c3", "c5") else cfg)
    if cfg == "c1":
        return scenes.cornell_box(xres, yres)
    if cfg in ("c3", "c5"):
        if key not in _SCENE_CACHE:
            _SCENE_CACHE[key] = scenes.sphere_field_scene(xres, yres)
        sc = _SCENE_CACHE[key]
        sc.XRes, sc.YRes = xres, yres          # c5 is the c3 scene at 4K
        if hasattr(sc.camera, "Aspect"):
            sc.camera.Aspect = xres / yres
        return sc
    sc = scenes.heightfield_scene(xres, yres, nq=NQ, motion=(cfg == "c4"))
    if cfg == "c2t":
        m = sc.meshes[0]
        xz = m.Verts[0][:, [0, 2]]
        m.UV = ((xz - xz.min(0)) / (xz.max(0) - xz.min(0)) * 8.0).astype(np.float32)
        sc.textures = [scenes.Texture("ground.png", scenes._test_texture(1024, 1024, 21))]
        [s for s in sc.shaders if s.Name == m.Shader[0]][0].DiffuseColour = "ground.png"
    return sc


# ---- host placement: page-locked buffers on the GPU's own NUMA node ---------------------------------------------------
_ALL_CPUS = None
_GPU_CPUS = None


def bind_to_gpu_numa(index):
    """Run this process on the CPUs of the NUMA node the GPU hangs off (sysfs: /sys/bus/pci/devices/<bdf>/numa_node), so that the
    page-locked ray / hit / frame buffers (first touch) and the thread that drives the copies are local to the GPU's PCIe root.
    Eight ranks each moving ~70 GB/s through one socket's memory is what capped the 8-GPU host-buffer rate. Returns a dict for
    the JSON line; does nothing when the topology cannot be read or the node's CPUs are outside this process's affinity mask."""
    global _ALL_CPUS, _GPU_CPUS
    info = {"numa_node": None, "bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        _ALL_CPUS = os.sched_getaffinity(0)
        local = cpus & _ALL_CPUS
        if local and local != _ALL_CPUS:
            os.sched_setaffinity(0, local)
            _GPU_CPUS = local
            info.update(bound=True, cpus=len(local))
    except Exception as e:   # no NVML / sysfs view in this container: stay where we are
        info["note"] = "not bound: %s" % type(e).__name__
    return info


class all_host_cpus:
    """The CPU legs (oracle on every host thread) run with the process's full affinity mask."""
    def __enter__(self):
        if _GPU_CPUS:
            os.sched_setaffinity(0, _ALL_CPUS)

    def __exit__(self, *a):
        if _GPU_CPUS:
            os.sched_setaffinity(0, _GPU_CPUS)


def make_oracle(scene):
    from oracle.binding import Oracle
    return Oracle(scene, motion_ref_compat=False)   # same MQBVH leaf mode as the device default ("fixed", DESIGN.md quirk b)


def cpu_frame(scene, table, spp, cores, target_s, faithful=False):
    """The oracle (reference-semantics CPU path) on a bounded number of iterations of the frame. faithful: the reference's own
    worker policy, min(10, cores) workers and global atomic ray counters."""
    with all_host_cpus():
        ora = make_oracle(scene)
        ora.set_scramble(table)
        nt = -cores if faithful else cores
        _, st1 = ora.render(0, 1, nthreads=nt)                                     # probe (also builds the trees)
        iters = int(min(spp, max(1, round(target_s / max(st1["seconds"], 1e-3)))))
        st = st1
        if iters > 1:
            ora.clear()
            _, st = ora.render(0, iters, nthreads=nt)
        else:
            iters = 1
    v = st["rays"] / st["seconds"] / 1e6
    workers = min(10, cores) if faithful else cores
    return {"value": v, "unit": "Mrays/s", "cores": workers, "per_core": v / workers, "kind": "port",
            "samples_per_s": scene.XRes * scene.YRes * iters / st["seconds"],
            "sample": "%d iteration(s) (spp) of the same %dx%d workload on %d host threads%s: %d rays in %.2f s" % (
                iters, scene.XRes, scene.YRes, workers, " (reference worker policy: min(10, cores) + global atomic counters)" if faithful else "",
                st["rays"], st["seconds"])}


# ---- the incoherent closest-hit batch of the headline --------------------------------------------------------------
def primary_rays(scene, cam, jx, jy):
    from vermeer_b200.host import RAY_DTYPE
    cam_m, ttf, asp = cam
    M = cam_m.reshape(4, 4).T
    xres, yres = scene.XRes, scene.YRes
    ys, xs = np.meshgrid(np.arange(yres), np.arange(xres), indexing="ij")
    sx = (-1 + 2 * (xs + jx) / xres).astype(np.float32)
    sy = -(-1 + 2 * (ys + jy) / yres).astype(np.float32)
    d = np.stack([sx * ttf, sy * (ttf / asp), -np.full_like(sx, scene.camera.Focal)], -1).reshape(-1, 3) @ M[:3, :3].T
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros(xres * yres, RAY_DTYPE)
    r["o"] = M[:3, 3]
    r["d"] = d.astype(np.float32)
    r["tmax"] = np.inf
    return r


def incoherent_batch(scene, cam, trace, seed0=5):
    """Four jittered primary passes (4 spp) -> ~5.9 M bounce rays: large enough that the persistent kernel's ramp-up and tail do
    not dominate the rate. `trace` = a closest-hit TraceProbe over a ray array (the GPU's or the oracle's: bit-identical)."""
    from vermeer_b200 import scenes
    parts = []
    for s, (jx, jy) in enumerate([(0.5, 0.5), (0.25, 0.75), (0.75, 0.25), (0.1, 0.4)]):
        prim = primary_rays(scene, cam, jx, jy)
        parts.append(scenes.incoherent_rays(prim, trace(prim), seed=seed0 + s))
    inc = np.concatenate(parts)
    return inc[np.random.default_rng(seed0).permutation(len(inc))]      # shuffled: no residual image-space coherence


def cpu_trace(scene, rays, cores, reps=1):
    with all_host_cpus():
        ora = make_oracle(scene)
        ora.trace(rays[:65536], nthreads=cores)
        t0 = time.perf_counter()
        for _ in range(reps):
            hits = ora.trace(rays, nthreads=cores)
        secs = (time.perf_counter() - t0) / reps
    return hits, len(rays) / secs / 1e6, secs


def time_batch(dev, torch, rays, steps, warmup, compact_e2e=True):
    """Device-resident and host-buffer rates of one closest-hit pass over `rays`. Returns a dict."""
    from vermeer_b200.host import HIT_DTYPE, HITC_DTYPE, RAY_DTYPE, RAYPD_DTYPE
    n = len(rays)
    d_r = torch.from_numpy(rays.view(np.uint8).reshape(n, 32)).cuda()
    d_h = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    for _ in range(max(3, warmup)):
        dev.trace_device(d_r.data_ptr(), n, d_h.data_ptr(), False)
    dev.reset_stats()
    ms = []
    for _ in range(steps):
        dev.trace_device(d_r.data_ptr(), n, d_h.data_ptr(), False)
        ms.append(dev.stats()["trace_ms"])
    st = dev.stats()
    hits = d_h.cpu().numpy().reshape(-1).view(HIT_DTYPE)
    # e2e: rays and hits in page-locked host memory, H2D + traversal + D2H inside the timed call
    pin_r = torch.from_numpy(rays.view(np.uint8).reshape(n, 32)).pin_memory().numpy().reshape(-1).view(RAY_DTYPE)
    pin_h = torch.empty((n, 32), dtype=torch.uint8, pin_memory=True).numpy().reshape(-1).view(HIT_DTYPE)
    pin_c = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True).numpy().reshape(-1).view(HITC_DTYPE)
    # 24-byte {P, D} records (VG_TRACE_RAYS_PD) when every ray of the batch is Ray.Init(P, D, +Inf) at Time 0: the same hits (checked below)
    pd_ok = bool(np.isposinf(rays["tmax"]).all() and (rays["time"] == 0).all())
    pin_pd = None
    if pd_ok:
        pd = np.zeros(n, RAYPD_DTYPE)
        pd["o"], pd["d"] = rays["o"], rays["d"]
        pin_pd = torch.from_numpy(pd.view(np.uint8).reshape(n, 24)).pin_memory().numpy().reshape(-1).view(RAYPD_DTYPE)
    e2e = {}
    pin_cpd = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True).numpy().reshape(-1).view(HITC_DTYPE) if pd_ok else None
    for name, inp, out, compact in (("full", pin_r, pin_h, False), ("compact", pin_r, pin_c, True), ("pd_compact", pin_pd, pin_cpd, True)):
        if (compact and not compact_e2e) or inp is None:
            continue
        try:
            dev.trace(inp, out=out, compact=compact)
        except RuntimeError:
            continue     # scenes the compact record does not cover
        t0 = time.perf_counter()
        reps = max(3, min(steps, 10))
        for _ in range(reps):
            dev.trace(inp, out=out, compact=compact)
        e2e[name] = (time.perf_counter() - t0) / reps * 1e3
    if "pd_compact" in e2e and pin_cpd.tobytes() != pin_c.tobytes():
        raise SystemExit("bench: VG_TRACE_RAYS_PD hits differ from the VgRay hits")
    res = {"rays": n, "ms_per_step": float(np.mean(ms)), "ms_best": float(np.min(ms)), "device_ms_total": float(np.sum(ms)),
           "hit_fraction": float((hits["prim"] >= 0).mean()),
           "nodesT_per_ray": st["nodes_t"] / max(1, steps) / n, "trisT_per_ray": st["tris_t"] / max(1, steps) / n,
           "e2e_ms": e2e, "hits": hits, "compact_hits": pin_c if "compact" in e2e else None}
    return res


# ---- one frame config ---------------------------------------------------------------------------------------------
class Rig:
    pass


def setup_frame(cfg, rank, world, local_rank, torch, dev=None):
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    from vermeer_b200.multigpu import init_library_comm
    r = Rig()
    r.cfg = cfg
    r.xres, r.yres, r.spp = cfg_res(cfg)
    r.scene = build_scene(cfg)
    t0 = time.time()
    r.host = HostScene(r.scene).prerender()
    r.prerender_s = time.time() - t0
    r.dev = dev if dev is not None else Device(local_rank)
    for k, v in [kv.split("=") for kv in os.environ.get("VG_OPTIONS", "").split(",") if kv]:   # options that act at upload (node_order)
        r.dev.set_option(k, int(v))
    t0 = time.time()
    r.dev.upload(r.host)
    r.upload_s = time.time() - t0
    if world > 1:
        if not getattr(r.dev, "_comm_ready", False):
            init_library_comm(r.dev, rank, world)
            r.dev._comm_ready = True
        else:
            r.dev.set_partition(rank, world)
    # the step's host-side input (per-pixel scramble table) and output (frame) live in page-locked host memory, as the bench
    # contract asks; the library DMAs from/to such buffers directly
    r.table_t = torch.from_numpy(scenes.splitmix64_table(SCRAMBLE_SEED, r.xres * r.yres).view(np.int64)).pin_memory()
    r.table = r.table_t.numpy().view(np.uint64)
    r.frame_t = torch.empty((r.yres, r.xres, 3), dtype=torch.float32, pin_memory=True)
    r.frame = r.frame_t.numpy()
    r.dev.set_scramble(r.table)
    # ~66 M paths per GPU and batch whatever the frame size and world
    r.iters_per_batch = int(min(r.spp, 64, max(1, round(ITERS_PER_BATCH * world * (1920 * 1080) / (r.xres * r.yres)))))
    r.dev.set_option("iters_per_batch", r.iters_per_batch)
    for k, v in [kv.split("=") for kv in os.environ.get("VG_OPTIONS", "").split(",") if kv]:   # tuning experiments, e.g. VG_OPTIONS=traversal=0
        r.dev.set_option(k, int(v))
    return r


def frame_leg(r, steps, warmup, rank, world, torch, dist):
    """K timed steps with the inputs resident in HBM (device time, max over ranks), then K steps end to end with host buffers."""
    dev, spp = r.dev, r.spp

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        dev.clear()
        dev.render(0, spp, fetch=False)
    dev.reset_stats()
    barrier()
    t0 = time.perf_counter()
    acc = dict(render_ms=0.0, closest_ms=0.0, shadow_ms=0.0, shade_ms=0.0, closest_launches=0, shadow_launches=0)
    for _ in range(steps):
        dev.clear()
        dev.render(0, spp, fetch=False)
        st = dev.stats()
        for k in acc:
            acc[k] += st[k]
    barrier()
    wall = time.perf_counter() - t0
    st = dev.stats()
    tvec = torch.tensor([acc["render_ms"], wall * 1e3], dtype=torch.float64, device="cuda")
    rvec = torch.tensor([float(st["rays"]), float(st["shadow_rays"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(rvec, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms = float(tvec[0]), float(tvec[1])
    rays_total, shadow_total = float(rvec[0]), float(rvec[1])

    # ---- e2e: host buffers every step; N > 1: the library's gather (NCCL) brings the owned pixels to rank 0 ---------
    def step_e2e():
        # vg_render_frame: H2D of this rank's rows of the scramble table from the caller's host buffer, clear, the render loop, the
        # library's exchange (N > 1: pack + ncclSend/Recv + scatter) and the D2H of the complete frame into the host buffer on rank 0,
        # pipelined in slices of tile rows (the copies and the exchange of one slice overlap the rendering of the next)
        if os.environ.get("VG_BENCH_E2E_PLAIN"):               # the three separate calls, for the A/B
            dev.set_scramble(r.table)
            dev.clear()
            if world > 1:
                dev.render(0, spp, fetch=False)
                dev.gather_frame(r.frame if rank == 0 else None)
            else:
                dev.render(0, spp, out=r.frame)
        else:
            dev.render_frame(r.table, 0, spp, out=r.frame if rank == 0 else None, clear=True)

    for _ in range(2):
        step_e2e()
    dev.reset_stats()
    barrier()
    t0 = time.perf_counter()
    gather_ms = 0.0
    for _ in range(steps):
        step_e2e()
        if world > 1:
            gather_ms += dev.stats()["gather_ms"]
    barrier()
    e2e_wall = time.perf_counter() - t0
    st2 = dev.stats()
    evec = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    r2 = torch.tensor([float(st2["rays"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(evec, op=dist.ReduceOp.MAX)
        dist.all_reduce(r2, op=dist.ReduceOp.SUM)
    e2e_s = float(evec[0])
    npix = r.xres * r.yres
    # the same frame with the shading trigonometry through float64 exactly like math/sincos.go (option precise_trig 1), N = 1 only
    precise = None
    if world == 1:
        dev.set_option("precise_trig", 1)
        dev.clear()
        dev.render(0, spp, fetch=False)
        dev.reset_stats()
        dev.clear()
        dev.render(0, spp, fetch=False)
        sp = dev.stats()
        precise = {"ms_per_step": sp["render_ms"], "shading_ms": sp["shade_ms"], "note": "precise_trig=1: float64 trig in shading (both paths are inside the 1e-3 RMSE bar, tested)"}
        dev.set_option("precise_trig", 0)
    out = {
        "value": rays_total / (dev_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": dev_ms / steps, "wall_ms_per_step": wall_ms / steps,
        "samples_per_s": npix * spp * steps / (dev_ms * 1e-3), "steps": steps,
        "rays_per_step": rays_total / steps, "shadow_rays_per_step": shadow_total / steps,
        "stage_ms_per_step": {"closest_traversal": acc["closest_ms"] / steps, "shadow_traversal": acc["shadow_ms"] / steps, "shading": acc["shade_ms"] / steps,
                              "raygen_resolve_accumulate": (acc["render_ms"] - acc["closest_ms"] - acc["shadow_ms"] - acc["shade_ms"]) / steps},
        "e2e": {"value": float(r2[0]) / e2e_s / 1e6, "unit": "Mrays/s", "samples_per_s": npix * spp * steps / e2e_s, "ms_per_step": e2e_s * 1e3 / steps,
                "h2d_bytes_per_step": (npix // world) * 48, "d2h_bytes_per_step": npix * 12,
                "gather_ms_per_step": gather_ms / steps if world > 1 else None},
        "scaling": "strong", "iters_per_batch": r.iters_per_batch, "trig": "fast (float32 libm; precise_trig=0)", "precise_trig": precise,
        "gpu_launches": int(st["kernel_launches"]),
        "shadow_level0_kernel": {0: "cooperative (k_trace_queue<1,3>)", 1: "per-lane, unordered (k_trace_queue<1,4>)"}.get(int(st.get("shadow_level0_kernel", -1)), "n/a") + "; measured choice",
        "_rank": dict(st=st, acc=acc, steps=steps),
    }
    return out


def gathered_frame_check(r, rank, world, local_rank, iters=4):
    """N > 1: the frame the library gathers on rank 0 (vg_gather_frame) must be the single-GPU render of the same iterations, bit for
    bit (ownership is disjoint, nothing is summed). Rank 0 renders the reference frame in a second, unpartitioned context."""
    from vermeer_b200.host import Device
    dev = r.dev
    dev.clear()
    dev.render(0, iters, fetch=False)
    dev.gather_frame(r.frame if rank == 0 else None)
    if rank != 0:
        return None
    got = r.frame.copy()
    solo = Device(local_rank).upload(r.host)
    solo.set_scramble(r.table)
    solo.set_option("iters_per_batch", iters)
    want = solo.render(0, iters)
    solo.close()
    same = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
    if not same:
        raise SystemExit("bench.py: the NCCL-gathered %d-GPU frame differs from the single-GPU frame (%d pixels)" % (world, int((got != want).any(-1).sum())))
    return same


def stage_roofline(cfg, frame, peaks, hbm_peak, hbm_src, ncu):
    """Per stage: live time per launch x traffic per ray from the ncu capture of this build -> achieved GB/s at each memory
    level against the measured peak of that level; plus the issue x SIMD ceiling of the capture."""
    st, acc, steps = frame["_rank"]["st"], frame["_rank"]["acc"], frame["_rank"]["steps"]
    cap = ncu.get("c3" if cfg in ("c3", "c5") else cfg, {})
    closest_rays = st["rays"] - st["shadow_rays"]
    bn, bt = (232.0, 84.0) if cfg == "c4" else (128.0, 48.0)   # algorithmic bytes per node visit / triangle test
    stages = {}
    spec = [("shadow", "shadow_ms", st["shadow_rays"], st["shadow_nodes_t"], st["shadow_tris_t"], acc["shadow_launches"]),
            ("closest", "closest_ms", closest_rays, st["nodes_t"], st["tris_t"], acc["closest_launches"]),
            ("shade", "shade_ms", closest_rays, 0, 0, acc["closest_launches"])]
    for name, key, rays, nodes_t, tris_t, launches in spec:
        ms = acc[key]
        if ms <= 0 or rays <= 0:
            continue
        launches = max(1, launches)
        e = {"ms_per_launch": ms / launches, "units_per_launch": rays / launches, "unit_name": "rays" if name != "shade" else "path vertices",
             "share_of_step": ms / acc["render_ms"]}
        if name != "shade":
            alg = 64.0 * rays + bn * nodes_t + bt * tris_t
            e["nodesT_per_ray"] = nodes_t / rays
            e["trisT_per_ray"] = tris_t / rays
            e["algorithmic"] = {"bytes_per_ray": alg / rays, "gbs": alg / (ms * 1e-3) / 1e9,
                                "note": "SURVEY.md 8d bytes model (64 + %g*NodesT + %g*TrisT per ray): mostly served by L1/L2, NOT memory traffic" % (bn, bt)}
        c = cap.get(name)
        if c:
            per = 1.0 / c["units"]
            secs = ms * 1e-3
            lv = {}
            for lvl, bytes_key, peak, src in (("hbm", "dram_bytes", hbm_peak, hbm_src),
                                              ("l2", "lts_bytes", peaks.get("l2_read_gbs"), "measured in this run (vg_measure_peaks: 48 MB re-read, ld.global.cg)"),
                                              ("l1", "l1tex_bytes", peaks.get("l1_read_gbs"), "measured in this run (vg_measure_peaks: 64 KB per SM re-read, ld.global.ca)")):
                if peak and c.get(bytes_key) is not None:
                    b = c[bytes_key] * per * rays
                    lv[lvl] = {"bytes_per_unit": c[bytes_key] * per, "achieved": b / secs / 1e9, "peak": peak, "unit": "GB/s", "frac": b / secs / 1e9 / peak,
                               "traffic_per_launch": b / launches, "peak_source": src}
            if c.get("l1tex_wavefronts_pct"):
                # the L1TEX data pipe counts wavefronts (one per 128-B line a warp's load touches), not bytes: a divergent warp
                # needs 32 of them for 32 x 32 B. Utilisation of that unit from the capture, scaled by live vs captured time per unit
                wf = c["l1tex_wavefronts_pct"] / 100.0 * (c["ms"] / c["units"]) / (ms / rays)
                lv["l1tex_wavefronts"] = {"frac": wf, "unit": "fraction of the L1TEX data pipe's sustained wavefront rate",
                                          "capture_pct": c["l1tex_wavefronts_pct"], "peak_source": "ncu: l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"}
            e["levels"] = lv
            e["issue"] = {"issue_slots_active": c["issue_active_pct"] / 100.0, "lanes_per_instruction": c["lanes_per_inst"],
                          "frac": c["issue_active_pct"] / 100.0 * c["lanes_per_inst"] / 32.0, "registers": c.get("registers"),
                          "warps_active_pct": c.get("warps_active_pct"), "thread_instructions_per_unit": c.get("thread_inst", 0) * per,
                          "note": "share of the SMs' lane-issue capacity doing work (issue slots active x active lanes / 32), from the ncu capture"}
            e["capture"] = {"kernel": c.get("kernel"), "ms": c.get("ms"), "units": c["units"], "file": cap.get("_file")}
            best = max(lv.items(), key=lambda kv: kv[1]["frac"]) if lv else None
            e["frac"] = max(best[1]["frac"] if best else 0.0, e["issue"]["frac"])
            e["binding"] = ("issue x SIMD lanes" if not best or e["issue"]["frac"] >= best[1]["frac"] else best[0])
        stages[name] = e
    return stages


def headline_roofline(stages, hbm_peak, hbm_src):
    """The contract's top-level object, for the stage that takes the largest share of the step."""
    if not stages:
        return None
    name, e = max(stages.items(), key=lambda kv: kv[1]["share_of_step"])
    lv = e.get("levels") or {}
    if lv:
        lvl, b = max(lv.items(), key=lambda kv: kv[1]["frac"])
        if lvl == "l1tex_wavefronts":   # the binding unit counts wavefronts; quote the byte figures of the L1 level beside the fraction
            l1 = lv.get("l1", {})
            out = {"bound": "l1tex (wavefronts)", "achieved": l1.get("achieved"), "peak": l1.get("peak"), "unit": "GB/s", "frac": b["frac"],
                   "traffic": l1.get("traffic_per_launch"), "peak_source": b["peak_source"],
                   "note": "frac = utilisation of the L1TEX data pipe (wavefronts/clk), the busiest unit; achieved/peak are the L1 BYTE rates "
                           "(measured peak for fully coalesced 128-bit loads), whose ratio is lower because divergent lanes use 32 B of every 128-B wavefront"}
        else:
            out = {"bound": {"hbm": "hbm", "l2": "l2", "l1": "l1tex"}[lvl], "achieved": b["achieved"], "peak": b["peak"], "unit": "GB/s", "frac": b["frac"],
                   "traffic": b["traffic_per_launch"], "peak_source": b["peak_source"]}
        out["levels"] = {k: v["frac"] for k, v in lv.items()}
        if "hbm" in lv:
            out["hbm"] = lv["hbm"]
    else:
        a = e.get("algorithmic", {"gbs": 0.0})
        out = {"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": hbm_src,
               "note": "no ncu capture of this config is committed: only the algorithmic-bytes figure is available (%.0f GB/s, not traffic)" % a["gbs"]}
    out["stage"] = name
    out["kernel"] = (e.get("capture") or {}).get("kernel")
    out["ms_per_launch"] = e["ms_per_launch"]
    out["share_of_step"] = e["share_of_step"]
    out["issue"] = e.get("issue")
    out["binding"] = e.get("binding")
    return out


def run_config(cfg, args, rank, world, local_rank, torch, dist, peaks, hbm, ncu, cores, dev=None, steps=None):
    xres, yres, spp = cfg_res(cfg)
    r = setup_frame(cfg, rank, world, local_rank, torch, dev=dev)
    # a step of the big configs lasts ~1 s (c3) / ~2 s (c5 on 8 GPUs): fewer repetitions keep the run within minutes
    k = steps or (args.steps if cfg in ("c1", "c2", "c4", "c2t") else max(2, min(args.steps, 3)))
    w = max(3, args.warmup)   # >= 3 also for the long configs: the library settles its measured choices (level-0 shadow kernel) in the first three calls
    f = frame_leg(r, k, w, rank, world, torch, dist)
    check = gathered_frame_check(r, rank, world, local_rank) if (world > 1 and cfg in ("c1", "c2")) else None
    out = None
    if rank == 0:
        stages = stage_roofline(cfg, f, peaks, hbm[0], hbm[1], ncu)
        out = {k2: v for k2, v in f.items() if not k2.startswith("_")}
        out["workload"] = CONFIGS[cfg]["workload"]
        out["triangles"] = r.scene.num_tris
        out["n_gpus"] = world
        out["stages"] = stages
        out["roofline"] = headline_roofline(stages, hbm[0], hbm[1])
        out["host_prerender_s"] = r.prerender_s
        if check is not None:
            out["gathered_frame_bit_identical_to_single_gpu"] = check
        out["upload_s"] = r.upload_s
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_frame(r.scene, r.table, spp, cores, 12.0 if cfg == "c2" else 4.0)
            out["cpu_baseline_faithful"] = cpu_frame(r.scene, r.table, spp, cores, 3.0, faithful=True)
            out["speedup_e2e_vs_cpu_all_cores"] = out["e2e"]["value"] / out["cpu_baseline"]["value"]
    return r, out


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from vermeer_b200.build import build
    from vermeer_b200.host import Device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the GPU path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not os.environ.get("VG_SO_PATH"):
        build()
    cores = os.cpu_count() or 1
    mp, hbm_src = measured_peaks()
    hbm = (float(mp["hbm_gbs"]), hbm_src)
    ncu = ncu_metrics()
    want = [c for c in (args.configs.split(",") if args.configs else ["c1", "c2", "c3", "c4"] + (["c5"] if world == 8 else [])) if c in CONFIGS]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    placement = bind_to_gpu_numa(local_rank)
    dev = Device(local_rank)
    peaks = dev.measure_peaks()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- C2: the frame (samples/s) and the headline incoherent closest-hit batch on the same scene ---------------------
    rig, c2 = run_config("c2", args, rank, world, local_rank, torch, dist, peaks, hbm, ncu, cores, dev=dev)
    scene, host = rig.scene, rig.host
    inc = incoherent_batch(scene, host.camera(), lambda rays: dev.trace(rays), seed0=5 + 16 * rank)
    barrier()
    t0 = time.perf_counter()
    tb = time_batch(dev, torch, inc, args.steps, args.warmup)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    n = tb["rays"]
    e2e_mode = "pd_compact" if "pd_compact" in tb["e2e_ms"] else ("compact" if "compact" in tb["e2e_ms"] else "full")
    vec = torch.tensor([tb["device_ms_total"], tb["e2e_ms"][e2e_mode], tb["e2e_ms"]["full"], tb["e2e_ms"].get("compact", tb["e2e_ms"]["full"])], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(n)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms_total, e2e_ms, e2e_full_ms, e2e_compact_ms = float(vec[0]), float(vec[1]), float(vec[2]), float(vec[3])
    rays_all = float(cnt[0])
    value = rays_all * args.steps / (dev_ms_total * 1e-3) / 1e6
    e2e_value = rays_all / (e2e_ms * 1e-3) / 1e6

    # CPU on the same rays + bit-identity, while the C2 scene is still the one uploaded (the slot table belongs to it)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        oh, cpu_v, secs = cpu_trace(scene, inc, cores, reps=3)
        same = bool(np.array_equal(oh["prim"], tb["hits"]["prim"]) and np.array_equal(oh["geom"], tb["hits"]["geom"]) and
                    np.array_equal(oh["t"].view(np.uint32), tb["hits"]["t"].view(np.uint32)) and
                    np.array_equal(oh["nodesT"], tb["hits"]["nodesT"]) and np.array_equal(oh["trisT"], tb["hits"]["trisT"]))
        compact_ok = None
        if tb["compact_hits"] is not None:
            prim_of, geom_of = dev.slot_table()
            ch = tb["compact_hits"]
            hit = ch["slot"] >= 0
            compact_ok = bool(np.array_equal(hit, oh["prim"] >= 0) and np.array_equal(prim_of[ch["slot"][hit]], oh["prim"][hit]) and
                              np.array_equal(geom_of[ch["slot"][hit]], oh["geom"][hit]) and np.array_equal(ch["t"].view(np.uint32), oh["t"].view(np.uint32)) and
                              np.array_equal(ch["u"].view(np.uint32), oh["u"].view(np.uint32)))
        cpu = {"value": cpu_v, "unit": "Mrays/s", "cores": cores, "per_core": cpu_v / cores, "kind": "port",
               "sample": "the same %d rays, 3 passes on all %d host threads, %.2f s per pass" % (n, cores, secs),
               "bit_identical_to_gpu": same, "compact_hits_identical": compact_ok}


    # parity-mode hits of the headline batch by original face, for the non-parity leg's check (the C2 scene is still uploaded)
    nonparity = None
    if rank == 0 and world == 1 and not args.no_nonparity:
        _, aidx16 = host.mesh_idxp(0)
        h16 = tb["hits"]
        face16 = np.where((h16["prim"] >= 0) & (h16["geom"] == 0), aidx16[np.maximum(h16["prim"], 0) % len(aidx16)], -1 - h16["geom"])
        nonparity = nonparity_leg(scene, inc, h16, face16, torch, local_rank)

    configs = {"c2": c2}
    for cfg in want:
        if cfg == "c2":
            continue
        if cfg == "c3" and "c5" in want and world == 8:
            pass
        rig_c, configs[cfg] = run_config(cfg, args, rank, world, local_rank, torch, dist, peaks, hbm, ncu, cores, dev=dev)
        if cfg == "c3" and rank == 0 and world == 1 and not args.no_wavefront:
            configs[cfg]["incoherent_wavefront"] = wavefront_leg(dev, torch, cores, args)
            configs[cfg]["incoherent_diffuse"] = diffuse_leg(dev, rig_c, torch, cores, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- headline bookkeeping on rank 0 ---------------------------------------------------------------------------
    alg_bytes = 64.0 + 128.0 * tb["nodesT_per_ray"] + 48.0 * tb["trisT_per_ray"]
    cap = ncu.get("c2_incoherent", {}).get("closest")
    roof = {"kernel": "k_trace_batch<0,2> (closest-hit QBVH traversal, cooperative leaf phase)", "stage": "closest-hit traversal of the incoherent batch",
            "ms_per_launch": tb["ms_per_step"], "units_per_launch": n, "nodesT_per_ray": tb["nodesT_per_ray"], "trisT_per_ray": tb["trisT_per_ray"],
            "algorithmic": {"bytes_per_ray": alg_bytes, "gbs": alg_bytes * n / (tb["ms_per_step"] * 1e-3) / 1e9,
                            "note": "SURVEY.md 8d bytes model: served by L1/L2 (the 54 MB scene is L2-resident), NOT memory traffic"}}
    if cap:
        per, secs = 1.0 / cap["units"], tb["ms_per_step"] * 1e-3
        lv = {}
        for lvl, key, peak, src in (("hbm", "dram_bytes", hbm[0], hbm[1]), ("l2", "lts_bytes", peaks["l2_read_gbs"], "measured in this run (vg_measure_peaks)"),
                                    ("l1", "l1tex_bytes", peaks["l1_read_gbs"], "measured in this run (vg_measure_peaks)")):
            b = cap[key] * per * n
            lv[lvl] = {"bytes_per_unit": cap[key] * per, "achieved": b / secs / 1e9, "peak": peak, "unit": "GB/s", "frac": b / secs / 1e9 / peak, "traffic_per_launch": b, "peak_source": src}
        lvl, b = max(lv.items(), key=lambda kv: kv[1]["frac"])
        wf = cap.get("l1tex_wavefronts_pct")
        if wf:
            wfrac = wf / 100.0 * (cap["ms"] / cap["units"]) / (tb["ms_per_step"] / n)
            lv["l1tex_wavefronts"] = {"frac": wfrac, "capture_pct": wf, "unit": "fraction of the L1TEX data pipe's sustained wavefront rate",
                                      "peak_source": "ncu: l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"}
            if wfrac > b["frac"]:
                lvl, b = "l1", dict(lv["l1"], frac=wfrac, peak_source=lv["l1tex_wavefronts"]["peak_source"])
                roof["note"] = ("frac = utilisation of the L1TEX data pipe (wavefronts/clk), the busiest unit of this kernel; achieved/peak are the L1 byte "
                                "rates, whose ratio (levels.l1.frac) is lower because a divergent warp uses 32 B of every 128-B wavefront")
        roof.update({"bound": {"hbm": "hbm", "l2": "l2", "l1": "l1tex"}[lvl], "achieved": b["achieved"], "peak": b["peak"], "unit": "GB/s", "frac": b["frac"],
                     "traffic": b["traffic_per_launch"], "peak_source": b["peak_source"], "levels": lv,
                     "issue": {"issue_slots_active": cap["issue_active_pct"] / 100.0, "lanes_per_instruction": cap["lanes_per_inst"],
                               "frac": cap["issue_active_pct"] / 100.0 * cap["lanes_per_inst"] / 32.0},
                     "capture": {"kernel": cap.get("kernel"), "ms": cap.get("ms"), "units": cap["units"], "file": ncu.get("c2_incoherent", {}).get("_file")}})
    else:
        roof.update({"bound": "hbm", "achieved": None, "peak": hbm[0], "unit": "GB/s", "frac": None, "traffic": None,
                     "note": "no committed ncu capture for this kernel"})

    out = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": HEADLINE_WORKLOAD, "rays_per_step_per_gpu": n, "triangles": scene.num_tris, "hit_fraction": tb["hit_fraction"],
                   "partition": "every rank traces its own batch (independent rays, no collective)" if world > 1 else "single GPU",
                   "l2_policy": "ray + hit streams of a step (%d MB) exceed the 126 MB L2; the 54 MB scene is L2-resident by nature of the config" % (n * 64 // 1000000)},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": n * (24 if e2e_mode == "pd_compact" else 32), "d2h_bytes_per_step": n * (32 if e2e_mode == "full" else 16),
                "ms_per_step": e2e_ms, "mode": "vg_trace_batch, page-locked host buffers, " +
                {"pd_compact": "VG_TRACE_RAYS_PD (24-byte {P, D} rays = Ray.Init(P, D, +Inf)) | VG_TRACE_COMPACT_HITS (16-byte hits); hits byte-identical to the VgRay call's",
                 "compact": "32-byte VgRay, VG_TRACE_COMPACT_HITS (16-byte hits)", "full": "32-byte VgRay, 32-byte VgHit"}[e2e_mode],
                "all_modes_ms_rank0": tb["e2e_ms"],
                "vgray_compact": {"value": rays_all / (e2e_compact_ms * 1e-3) / 1e6, "ms_per_step": e2e_compact_ms, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 16},
                "full_vghit": {"value": rays_all / (e2e_full_ms * 1e-3) / 1e6, "ms_per_step": e2e_full_ms, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32}},
        "gpu_launches": args.steps, "wall_s_timed_region": wall, "host_placement": placement,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "samples_per_s": c2["samples_per_s"], "samples_per_s_e2e": c2["e2e"]["samples_per_s"],
        "frame": {k: v for k, v in c2.items() if k not in ("stages",)},
        "configs": configs,
        "incoherent_wavefront": (configs.get("c3") or {}).get("incoherent_wavefront"),
        "incoherent_diffuse": (configs.get("c3") or {}).get("incoherent_diffuse"),
        "nonparity": nonparity,
        "measured_peaks": {"hbm_gbs": hbm[0], "hbm_source": hbm[1], **peaks},
        "speedup_vs_cpu": None if not cpu else {"device_resident": value / cpu["value"], "e2e": e2e_value / cpu["value"], "cores": cores,
                                                "frame_e2e": c2.get("speedup_e2e_vs_cpu_all_cores")},
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def nonparity_leg(scene, inc, ref_hits, ref_face, torch, local_rank, leaf_max=4):
    """Opt-in NON-PARITY mode (vh_set_option "leaf_max"): the reference's own builder run with leafMax 4 instead of
    16 — another tree over the same triangles, same intersection routine. Reported beside the parity numbers, never instead of
    them. Same frame (C2, 64 spp) and same incoherent batch; the hits must be the same t and the same face."""
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    host = HostScene(scene, leaf_max=leaf_max).prerender()
    dev = Device(local_rank).upload(host)
    dev.set_scramble(scenes.splitmix64_table(SCRAMBLE_SEED, scene.XRes * scene.YRes))
    dev.set_option("iters_per_batch", ITERS_PER_BATCH)
    ms = []
    for i in range(4):
        dev.clear()
        dev.render(0, 64, fetch=False)
        st = dev.stats()
        if i >= 1:
            ms.append((st["render_ms"], st["closest_ms"], st["shadow_ms"], st["shade_ms"]))
    m = np.mean(np.asarray(ms), 0)
    dev.reset_stats()
    dev.clear()
    dev.render(0, 64, fetch=False)
    rays = dev.stats()["rays"]
    tb = time_batch(dev, torch, inc, 5, 3, compact_e2e=False)
    _, aidx = host.mesh_idxp(0)
    h = tb["hits"]
    face = np.where((h["prim"] >= 0) & (h["geom"] == 0), aidx[np.maximum(h["prim"], 0) % len(aidx)], -1 - h["geom"])
    out = {"mode": "leaf_max=%d (reference builder, smaller leaves; NOT the reference's tree)" % leaf_max, "nodes": host.mesh_info(0)["nodes"],
           "frame": {"value": rays / m[0] / 1e3, "unit": "Mrays/s", "ms_per_step": float(m[0]),
                     "stage_ms_per_step": {"closest_traversal": float(m[1]), "shadow_traversal": float(m[2]), "shading": float(m[3])}},
           "incoherent": {"value": tb["rays"] / tb["ms_per_step"] / 1e3, "unit": "Mrays/s", "nodesT_per_ray": tb["nodesT_per_ray"], "trisT_per_ray": tb["trisT_per_ray"]},
           "same_t_as_parity_mode": float((h["t"].view(np.uint32) == ref_hits["t"].view(np.uint32)).mean()),
           "same_face_as_parity_mode": float((face == ref_face).mean())}
    dev.close()
    return out


def wavefront_leg(dev, torch, cores, args):
    """Incoherent closest-hit rays taken from the wavefront itself: the level-1..3 mirror-bounce queues of the 10M-triangle scene
    (uploaded in `dev` by the c3 frame leg), two iterations' worth, shuffled. Not L2-resident (0.55 GB of nodes + triangles)."""
    dev.set_option("capture_levels", 0b1110)
    dev.clear()
    dev.render(0, 2, fetch=False)
    dev.set_option("capture_levels", 0)
    rays = dev.captured_rays()
    rays = rays[np.random.default_rng(11).permutation(len(rays))]
    tb = time_batch(dev, torch, rays, max(3, min(args.steps, 10)), 3)
    n = tb["rays"]
    res = {"workload": "C3-wavefront: level-1..3 mirror-bounce rays of the 10M-triangle sphere field, from vg_render's own ray queues (2 iterations), shuffled",
           "rays": n, "value": n / tb["ms_per_step"] / 1e3, "unit": "Mrays/s", "ms_per_step": tb["ms_per_step"], "hit_fraction": tb["hit_fraction"],
           "nodesT_per_ray": tb["nodesT_per_ray"], "trisT_per_ray": tb["trisT_per_ray"],
           "e2e": {k: n / v / 1e3 for k, v in tb["e2e_ms"].items()}}
    if not args.no_cpu:
        sample = rays[: min(n, 1 << 21)]
        oh, v, secs = cpu_trace(build_scene("c3"), sample, cores)
        g = tb["hits"][: len(sample)]
        res["cpu"] = {"value": v, "unit": "Mrays/s", "cores": cores, "per_core": v / cores, "kind": "port", "sample": "%d of the same rays, %.2f s" % (len(sample), secs),
                      "bit_identical_to_gpu": bool(np.array_equal(oh["prim"], g["prim"]) and np.array_equal(oh["geom"], g["geom"]) and
                                                   np.array_equal(oh["t"].view(np.uint32), g["t"].view(np.uint32)))}
        res["speedup_vs_cpu"] = {"device_resident": res["value"] / v, "e2e": {k: x / v for k, x in res["e2e"].items()}}
    return res


def diffuse_leg(dev, rig, torch, cores, args):
    """BASELINE.json's third config by name: 4-bounce incoherent DIFFUSE paths on the 10M-triangle scene, as a traversal workload (the
    reference's shader has no diffuse indirect — DESIGN.md quirk e — so the frame config C3 follows mirror chains; this leg traces what
    a diffuse path tracer would): one 1080p camera pass, then four cosine-hemisphere bounces off the hit points
    (scenes.diffuse_bounce_rays), every bounce's rays shuffled. Device-resident rate per bounce and for the four together; CPU on the
    same rays with bit-identity."""
    from vermeer_b200 import scenes
    scene, host = rig.scene, rig.host
    cur = primary_rays(scene, host.camera(), 0.5, 0.5)
    hits = dev.trace(cur)
    per, allr = [], []
    for b in range(1, 5):
        cur = scenes.diffuse_bounce_rays(scene, cur, hits, seed=40 + b)
        cur = cur[np.random.default_rng(50 + b).permutation(len(cur))]
        tb = time_batch(dev, torch, cur, 3, 3, compact_e2e=False)
        hits = tb["hits"]
        per.append({"bounce": b, "rays": len(cur), "value": len(cur) / tb["ms_per_step"] / 1e3, "hit_fraction": tb["hit_fraction"],
                    "nodesT_per_ray": tb["nodesT_per_ray"], "trisT_per_ray": tb["trisT_per_ray"]})
        allr.append(cur)
    rays = np.concatenate(allr)
    rays = rays[np.random.default_rng(60).permutation(len(rays))]
    tb = time_batch(dev, torch, rays, max(3, min(args.steps, 10)), 3)
    n = tb["rays"]
    res = {"workload": "C3-diffuse: bounces 1-4 of cosine-hemisphere diffuse paths on the 10M-triangle sphere field (one 1080p camera pass), shuffled",
           "rays": n, "value": n / tb["ms_per_step"] / 1e3, "unit": "Mrays/s", "ms_per_step": tb["ms_per_step"], "hit_fraction": tb["hit_fraction"],
           "nodesT_per_ray": tb["nodesT_per_ray"], "trisT_per_ray": tb["trisT_per_ray"], "per_bounce": per,
           "e2e": {k: n / v / 1e3 for k, v in tb["e2e_ms"].items()}}
    if not args.no_cpu:
        sample = rays[: min(n, 1 << 21)]
        oh, v, secs = cpu_trace(scene, sample, cores)
        g = tb["hits"][: len(sample)]
        res["cpu"] = {"value": v, "unit": "Mrays/s", "cores": cores, "per_core": v / cores, "kind": "port", "sample": "%d of the same rays, %.2f s" % (len(sample), secs),
                      "bit_identical_to_gpu": bool(np.array_equal(oh["prim"], g["prim"]) and np.array_equal(oh["geom"], g["geom"]) and
                                                   np.array_equal(oh["t"].view(np.uint32), g["t"].view(np.uint32)) and
                                                   np.array_equal(oh["nodesT"], g["nodesT"]) and np.array_equal(oh["trisT"], g["trisT"]))}
        res["speedup_vs_cpu"] = {"device_resident": res["value"] / v, "e2e": {k: x / v for k, x in res["e2e"].items()}}
    return res


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores. The Go reference cannot
    be compiled in this image (no Go toolchain), so this is the oracle port (kind "port"), all host threads; same metric and
    workload as the GPU arm: one step = one closest-hit pass over the incoherent batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vermeer_b200 import scenes
    cores = os.cpu_count() or 1
    scene = build_scene("c2")
    ora = make_oracle(scene)
    cam = ora.camera_matrix()
    inc = incoherent_batch(scene, cam, lambda rays: ora.trace(rays, nthreads=cores), seed0=5)
    n = len(inc)
    for _ in range(min(args.warmup, 2)):
        ora.trace(inc, nthreads=cores)
    secs = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ora.trace(inc, nthreads=cores)
        secs += time.perf_counter() - t0
    v = n * args.steps / secs / 1e6
    sample = "the whole batch (%d rays) per step on all %d host threads" % (n, cores)
    # the frame half of the metric: a bounded sample of the C2 frame
    table = scenes.splitmix64_table(SCRAMBLE_SEED, scene.XRes * scene.YRes)
    frame = cpu_frame(scene, table, 64, cores, 8.0)
    faithful = cpu_frame(scene, table, 64, cores, 4.0, faithful=True)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": HEADLINE_WORKLOAD, "rays_per_step_per_gpu": n, "triangles": scene.num_tris, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "per_core": v / cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "samples_per_s": frame["samples_per_s"], "frame": frame, "frame_faithful": faithful,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-wavefront", action="store_true", help="skip the C3 wavefront incoherent batch")
    ap.add_argument("--no-nonparity", action="store_true", help="skip the opt-in non-parity (leaf_max=4) leg")
    ap.add_argument("--configs", default=os.environ.get("VG_BENCH_CONFIGS", ""), help="comma list of frame configs (default c1,c2,c3,c4 and c5 at 8 GPUs); c2 always runs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

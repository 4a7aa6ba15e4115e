"""ctypes harness over libvermeer_b200.so (the C ABI of include/vermeer_gpu.h).

Used by tests/ and bench.py; numpy in, numpy out, no torch types in any signature.  There is no CPU
fallback here: `Device()` raises when the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .scenes import SceneDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("VG_SO_PATH", os.path.join(_HERE, "libvermeer_b200.so"))

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("tmax", np.float32), ("time", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("w", np.float32),
                      ("prim", np.int32), ("geom", np.int32), ("nodesT", np.int32), ("trisT", np.int32)])
NODE_DTYPE = np.dtype([("Boxes", np.float32, 24), ("Axis0", np.uint32), ("Axis1", np.uint32), ("Axis2", np.uint32),
                       ("Children", np.int32, 4), ("Parent", np.int32)])
MNODE_DTYPE = np.dtype([("Axis0", np.int32), ("Axis1", np.int32), ("Axis2", np.int32), ("Children", np.int32, 4),
                        ("Parent", np.int32), ("pad", np.uint32, 2)])

VG_TRACE_ANY_HIT = 1
VG_TRACE_COMPACT_HITS = 2
VG_TRACE_RAYS_PD = 4
RAYPD_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3)])   # VgRayPD: tmax = +Inf, time = 0
HITC_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("slot", np.int32)])

# every symbol include/vermeer_gpu.h declares (tests check the library exports all of them)
DECLARED_SYMBOLS = [
    "vg_create", "vg_destroy", "vg_last_error", "vg_device_count", "vg_scene_begin", "vg_mesh_upload", "vg_mesh_upload_motion",
    "vg_sphere_upload", "vg_instance_upload", "vg_scene_upload", "vg_scene_upload_motion", "vg_scene_commit", "vg_set_materials", "vg_set_lights", "vg_set_area_lights", "vg_set_camera", "vg_set_camera_motion", "vg_set_frame",
    "vg_set_partition", "vg_set_scramble", "vg_set_filter", "vg_set_option", "vg_trace_batch", "vg_trace_batch_device", "vg_render", "vg_clear_framebuffer",
    "vg_framebuffer_device", "vg_get_stats", "vg_reset_stats",
    "vg_comm_unique_id", "vg_comm_init", "vg_comm_destroy", "vg_gather_frame", "vg_render_frame", "vg_nccl_version", "vg_owned_pixels", "vg_measure_peaks", "vg_slot_table", "vg_captured_rays",
    "vg_texture_upload", "vg_textures_clear", "vg_texture_levels", "vg_texture_read_level", "vg_material_set_texture", "vg_mesh_set_uv", "vg_texture_sample_batch",
    "vh_add_texture", "vh_shader_set_texture", "vh_polymesh_set_uv", "vg_build_qbvh", "vg_build_qbvh_nodes", "vh_prerender_device",
    "vh_set_option", "vh_scene_create", "vh_scene_destroy", "vh_last_error", "vh_registered_nodes", "vh_set_globals", "vh_add_shader_std", "vh_add_shader_debug", "vh_add_polymesh",
    "vh_add_filter", "vh_add_instance", "vh_add_trilight", "vh_add_disklight", "vh_add_spherelight", "vh_parse_vnf", "vh_load_vnf", "vh_globals", "vh_postrender", "vh_rgbe", "vh_set_camera_lookat", "vh_set_camera_keys", "vh_camera_decomp", "vh_prerender", "vh_upload", "vh_num_geoms", "vh_scene_info", "vh_scene_nodes",
    "vh_scene_motion_nodes", "vh_scene_geom_order", "vh_mesh_info", "vh_mesh_nodes", "vh_mesh_motion_nodes", "vh_mesh_idxp", "vh_camera",
]


class VgMaterial(C.Structure):
    _fields_ = [("mask", C.c_uint32), ("emission_colour", C.c_float * 3), ("emission_strength", C.c_float),
                ("diffuse_colour", C.c_float * 3), ("diffuse_strength", C.c_float), ("diffuse_roughness", C.c_float),
                ("spec1_colour", C.c_float * 3), ("spec1_strength", C.c_float), ("spec1_roughness", C.c_float), ("ior", C.c_float),
                ("spec1_fresnel_model", C.c_int32), ("spec1_fresnel_refl", C.c_float * 3), ("spec1_fresnel_edge", C.c_float * 3)]


class VgCamera(C.Structure):
    _fields_ = [("local_to_world", C.c_float * 16), ("tan_theta_focal", C.c_float), ("aspect", C.c_float), ("focal", C.c_float), ("radius", C.c_float)]


class VgStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("nodes_t", C.c_uint64), ("tris_t", C.c_uint64),
                ("shadow_nodes_t", C.c_uint64), ("shadow_tris_t", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("closest_launches", C.c_uint64), ("shadow_launches", C.c_uint64),
                ("render_ms", C.c_double), ("trace_ms", C.c_double), ("closest_ms", C.c_double), ("shadow_ms", C.c_double),
                ("shade_ms", C.c_double), ("gather_ms", C.c_double), ("max_stack_depth", C.c_uint64), ("shadow_level0_kernel", C.c_int64)]


class VgPeaks(C.Structure):
    _fields_ = [("hbm_read_gbs", C.c_double), ("l2_read_gbs", C.c_double), ("l1_read_gbs", C.c_double),
                ("hbm_buffer_bytes", C.c_double), ("l2_buffer_bytes", C.c_double), ("l1_window_bytes", C.c_double),
                ("sm_count", C.c_int32), ("pad_", C.c_int32)]


VG_COMM_ID_BYTES = 128


_LIB = None


def load_library():
    """dlopen the in-tree library; raises if it has not been built (python -m vermeer_b200.build)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("libvermeer_b200.so is missing: run `python -m vermeer_b200.build` (there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        L.vg_last_error.restype = C.c_char_p
        L.vg_last_error.argtypes = [C.c_void_p]
        L.vh_last_error.restype = C.c_char_p
        L.vh_last_error.argtypes = [C.c_void_p]
        L.vg_destroy.argtypes = [C.c_void_p]
        L.vg_destroy.restype = None
        L.vh_scene_destroy.argtypes = [C.c_void_p]
        L.vh_scene_destroy.restype = None
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def material_struct(sh) -> VgMaterial:
    mask, p = sh.packed()
    m = VgMaterial()
    m.mask = mask
    m.emission_colour[:] = p[0:3]
    m.emission_strength = p[3]
    m.diffuse_colour[:] = p[4:7]
    m.diffuse_strength = p[7]
    m.diffuse_roughness = p[8]
    m.spec1_colour[:] = p[9:12]
    m.spec1_strength = p[12]
    m.spec1_roughness = p[13]
    m.ior = p[14]
    m.spec1_fresnel_model = int(p[15])
    m.spec1_fresnel_refl[:] = p[16:19]
    m.spec1_fresnel_edge[:] = p[19:22]
    return m


class HostScene:
    """vh_* layer: the reference's node registry + PreRender, from a SceneDesc."""

    def __init__(self, scene: SceneDesc, leaf_max: int | None = None):
        self.L = load_library()
        self.scene = scene
        h = C.c_void_p()
        if self.L.vh_scene_create(C.byref(h)) != 0:
            raise RuntimeError("vh_scene_create failed")
        self.h = h
        L = self.L
        if leaf_max is not None:   # opt-in non-parity mode: the reference's builder with another leafMax (vh_set_option)
            self._chk(L.vh_set_option(h, b"leaf_max", int(leaf_max)))
        self._chk(L.vh_set_globals(h, scene.XRes, scene.YRes, scene.MaxIter))
        for s in scene.shaders:
            if hasattr(s, "Colour"):   # scenes.DebugShader
                self._chk(L.vh_add_shader_debug(h, s.Name.encode(), _f3(s.Colour)))
                continue
            m = material_struct(s)
            self._chk(L.vh_add_shader_std(h, s.Name.encode(), C.byref(m)))
            from .scenes import SHADER_SLOTS
            for slot, (name, _, _) in enumerate(SHADER_SLOTS):
                v = getattr(s, name)
                if isinstance(v, str) and name != "Spec1FresnelModel":
                    self._chk(L.vh_shader_set_texture(h, s.Name.encode(), slot, v.encode()))
        for t in getattr(scene, "textures", []):
            self.add_texture(t)
        for m in scene.meshes:
            keys, nverts, _ = m.Verts.shape
            self._chk(L.vh_add_polymesh(
                h, m.Name.encode(), _p(m.Verts), nverts, keys,
                _p(m.PolyCount), 0 if m.PolyCount is None else len(m.PolyCount),
                _p(m.FaceIdx), 0 if m.FaceIdx is None else len(m.FaceIdx),
                "\n".join(m.Shader).encode(),
                _p(m.ShaderIdx), 0 if m.ShaderIdx is None else len(m.ShaderIdx),
                _p(m.Normals), 0 if m.Normals is None else len(m.Normals),
                _p(m.NormalIdx), 0 if m.NormalIdx is None else len(m.NormalIdx),
                C.c_float(m.RayBias)))
            if getattr(m, "UV", None) is not None:
                self._chk(L.vh_polymesh_set_uv(h, m.Name.encode(), _p(m.UV), len(m.UV), _p(m.UVIdx), 0 if m.UVIdx is None else len(m.UVIdx)))
        for ins in getattr(scene, "instances", []):
            bmin = np.ascontiguousarray(ins.BMin, np.float32).reshape(-1, 3)
            bmax = np.ascontiguousarray(ins.BMax, np.float32).reshape(-1, 3)
            tr = np.ascontiguousarray(ins.Transform, np.float32).reshape(-1, 16)
            self._chk(L.vh_add_instance(h, ins.Name.encode(), ins.Geom.encode(), _p(bmin), _p(bmax), len(bmin), _p(tr), len(tr)))
        for l in scene.lights:
            kind = type(l).__name__
            if kind == "TriLight":
                self._chk(L.vh_add_trilight(h, l.Name.encode(), _f3(l.P0), _f3(l.P1), _f3(l.P2), l.Shader.encode(), l.Samples))
            elif kind == "DiskLight":
                self._chk(L.vh_add_disklight(h, l.Name.encode(), _f3(l.P), _f3(l.LookAt), _f3(l.Up), C.c_float(l.Radius), l.Shader.encode(),
                                             int(l.Segments), int(l.Samples)))
            elif kind == "SphereLight":
                self._chk(L.vh_add_spherelight(h, l.Name.encode(), _f3(l.P), C.c_float(l.Radius), l.Shader.encode(), int(l.Samples)))
            else:
                raise ValueError("unknown light node %r" % kind)
        if scene.filter is not None:
            f = scene.filter
            self._chk(L.vh_add_filter(h, f.Type.encode(), f.Name.encode(), C.c_float(f.Width or 0), int(f.Res or 0), C.c_float(f.Peak or 0)))
        c = scene.camera
        if getattr(c, "has_keys", False):
            fr, to, ro, w2l = c.keys()
            self._chk(L.vh_set_camera_keys(h, c.Type.encode(), _p(fr), len(fr), _p(to), len(to), _p(ro), len(ro), _f3(c.Up), _p(w2l), len(w2l),
                                           C.c_float(c.Fov), C.c_float(c.Focal), C.c_float(c.Aspect), C.c_float(c.Radius)))
        else:
            self._chk(L.vh_set_camera_lookat(h, _f3(c.From), _f3(c.To), _f3(c.Up), C.c_float(c.Roll), C.c_float(c.Fov), C.c_float(c.Focal),
                                             C.c_float(c.Aspect), C.c_float(c.Radius)))

    @classmethod
    def from_vnf(cls, text: str | None = None, path: str | None = None, strict: bool = True):
        """nodes.Parse: build the scene from a .vnf description (text in memory or a file). With `strict`, any parse error the
        reference would have printed raises; otherwise the nodes that parsed are kept and `parse_errors`/`parse_log` tell."""
        from types import SimpleNamespace
        self = cls.__new__(cls)
        self.L = load_library()
        h = C.c_void_p()
        if self.L.vh_scene_create(C.byref(h)) != 0:
            raise RuntimeError("vh_scene_create failed")
        self.h = h
        if path is not None:
            n = self.L.vh_load_vnf(h, path.encode())
        else:
            b = (text or "").encode()
            n = self.L.vh_parse_vnf(h, b, C.c_size_t(len(b)), b"<memory>")
        self.parse_errors = n
        self.parse_log = self.L.vh_last_error(h).decode()
        if n < 0 or (strict and n > 0):
            raise RuntimeError("vnf (%d): %s" % (n, self.parse_log))
        g = np.zeros(3, np.int32)
        self._chk(self.L.vh_globals(h, _p(g)))
        self.scene = SimpleNamespace(XRes=int(g[0]), YRes=int(g[1]), MaxIter=int(g[2]))
        return self

    def add_texture(self, tex):
        """Register a decoded image (scenes.Texture) under the file name the shaders' texture maps use."""
        px = tex.rows_bottom_up()
        self._chk(self.L.vh_add_texture(self.h, tex.Name.encode(), int(px.shape[1]), int(px.shape[0]), _p(px)))

    def postrender(self, fb: np.ndarray):
        """core.PostRender: run the scene's OutputFloat / OutputHDR nodes on the finished frame."""
        fb = np.ascontiguousarray(fb, np.float32)
        self._chk(self.L.vh_postrender(self.h, _p(fb), int(fb.shape[1]), int(fb.shape[0])))

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("vermeer host (%d): %s" % (rc, self.L.vh_last_error(self.h).decode()))

    def prerender(self, device=None):
        """core.PreRender. `device` (a host.Device): build the static meshes' QBVHs on that GPU (vh_prerender_device)."""
        if device is not None:
            self._chk(self.L.vh_prerender_device(self.h, device.h))
        else:
            self._chk(self.L.vh_prerender(self.h))
        return self

    def close(self):
        if getattr(self, "h", None):
            self.L.vh_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # inspection --------------------------------------------------------------------------------
    def num_geoms(self):
        return self.L.vh_num_geoms(self.h)

    def scene_info(self):
        o = np.zeros(4, np.int32)
        self._chk(self.L.vh_scene_info(self.h, _p(o)))
        return dict(nodes=int(o[0]), motion=bool(o[1]), keys=int(o[2]), slots=int(o[3]))

    def scene_nodes(self):
        info = self.scene_info()
        if info["motion"]:
            topo = np.zeros(info["nodes"], MNODE_DTYPE)
            boxes = np.zeros((info["keys"], info["nodes"], 24), np.float32)
            self._chk(self.L.vh_scene_motion_nodes(self.h, _p(topo), _p(boxes)))
            return topo, boxes
        out = np.zeros(info["nodes"], NODE_DTYPE)
        self._chk(self.L.vh_scene_nodes(self.h, _p(out)))
        return out

    def scene_geom_order(self):
        o = np.zeros(self.num_geoms(), np.int32)
        self._chk(self.L.vh_scene_geom_order(self.h, _p(o)))
        return o

    def mesh_info(self, gid):
        o = np.zeros(6, np.int32)
        self._chk(self.L.vh_mesh_info(self.h, gid, _p(o)))
        return dict(nodes=int(o[0]), tris=int(o[1]), keys=int(o[2]), nverts=int(o[3]), motion=bool(o[4]), normals=bool(o[5]))

    def mesh_nodes(self, gid):
        info = self.mesh_info(gid)
        if info["motion"]:
            topo = np.zeros(info["nodes"], MNODE_DTYPE)
            boxes = np.zeros((info["keys"], info["nodes"], 24), np.float32)
            self._chk(self.L.vh_mesh_motion_nodes(self.h, gid, _p(topo), _p(boxes)))
            return topo, boxes
        out = np.zeros(info["nodes"], NODE_DTYPE)
        self._chk(self.L.vh_mesh_nodes(self.h, gid, _p(out)))
        return out

    def mesh_idxp(self, gid):
        info = self.mesh_info(gid)
        idxp = np.zeros(info["tris"] * 3, np.uint32)
        aidx = np.zeros(info["tris"], np.int32)
        self._chk(self.L.vh_mesh_idxp(self.h, gid, _p(idxp), _p(aidx)))
        return idxp, aidx

    def camera(self):
        c = VgCamera()
        self._chk(self.L.vh_camera(self.h, C.byref(c)))
        return np.asarray(c.local_to_world[:], np.float32), c.tan_theta_focal, c.aspect

    def camera_decomp(self):
        """Camera.decomp after PreRender: [keys, 23] float32 (T, R{X,Y,Z,W}, S column major)."""
        n = self.L.vh_camera_decomp(self.h, None)
        if n < 0:
            raise RuntimeError("vh_camera_decomp: %s" % self.L.vh_last_error(self.h).decode())
        out = np.zeros((n, 23), np.float32)
        if n:
            self.L.vh_camera_decomp(self.h, _p(out))
        return out


class Device:
    """vg_* layer: one context on one B200."""

    def __init__(self, ordinal: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.vg_create(C.byref(h), ordinal)
        if rc != 0:
            raise RuntimeError("vg_create failed (%d): %s" % (rc, self.L.vg_last_error(None).decode()))
        self.h = h
        self.xres = self.yres = 0

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("vermeer device (%d): %s" % (rc, self.L.vg_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.vg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, host: HostScene, motion_ref_compat: bool = False):
        rc = self.L.vh_upload(host.h, self.h, 1 if motion_ref_compat else 0)
        if rc != 0:
            raise RuntimeError("vh_upload (%d): %s" % (rc, self.L.vh_last_error(host.h).decode()))
        self.xres, self.yres = host.scene.XRes, host.scene.YRes
        return self

    def set_partition(self, rank: int, world: int):
        self._chk(self.L.vg_set_partition(self.h, rank, world))

    def set_scramble(self, table: np.ndarray):
        table = np.ascontiguousarray(table, np.uint64)
        self._chk(self.L.vg_set_scramble(self.h, _p(table), C.c_int64(table.shape[0])))

    def set_option(self, name: str, value: int):
        self._chk(self.L.vg_set_option(self.h, name.encode(), int(value)))

    def trace(self, rays: np.ndarray, any_hit: bool = False, out: np.ndarray | None = None, compact: bool = False) -> np.ndarray:
        """vg_trace_batch with HOST buffers (H2D + kernel + D2H inside the call). compact: 16-byte VgHitCompact records.
        A RAYPD_DTYPE array (24-byte {P, D} records: tmax = +Inf, time = 0) goes through as VG_TRACE_RAYS_PD."""
        pd = isinstance(rays, np.ndarray) and rays.dtype == RAYPD_DTYPE
        if not (isinstance(rays, np.ndarray) and rays.dtype in (RAY_DTYPE, RAYPD_DTYPE) and rays.flags.c_contiguous):
            rays = np.ascontiguousarray(rays, RAYPD_DTYPE if pd else RAY_DTYPE)      # (a page-locked array passes through untouched)
        hits = np.empty(len(rays), HITC_DTYPE if compact else HIT_DTYPE) if out is None else out
        flags = (VG_TRACE_ANY_HIT if any_hit else 0) | (VG_TRACE_COMPACT_HITS if compact else 0) | (VG_TRACE_RAYS_PD if pd else 0)
        self._chk(self.L.vg_trace_batch(self.h, _p(rays), C.c_int64(len(rays)), _p(hits), C.c_uint32(flags)))
        return hits

    def slot_table(self):
        """vg_slot_table: (prim_of_slot, geom_of_slot) for the static triangle slots compact hits refer to."""
        self.L.vg_slot_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        n = self.L.vg_slot_table(self.h, None, None, 0)
        if n < 0:
            self._chk(n)
        prim, geom = np.zeros(n, np.int32), np.zeros(n, np.int32)
        if self.L.vg_slot_table(self.h, _p(prim), _p(geom), n) < 0:
            self._chk(-1)
        return prim, geom

    def captured_rays(self) -> np.ndarray:
        """vg_captured_rays: the ray queues kept by the "capture_levels" option, as a VgRay array."""
        self.L.vg_captured_rays.restype = C.c_int64
        self.L.vg_captured_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        n = self.L.vg_captured_rays(self.h, None, 0)
        out = np.zeros(n, RAY_DTYPE)
        if n:
            self.L.vg_captured_rays(self.h, _p(out), n)
        return out

    def trace_device(self, d_rays_ptr: int, n: int, d_hits_ptr: int, any_hit: bool = False, pd: bool = False, compact: bool = False):
        """vg_trace_batch_device: rays/hits already resident in HBM (raw device pointers, e.g. torch .data_ptr())."""
        flags = (VG_TRACE_ANY_HIT if any_hit else 0) | (VG_TRACE_COMPACT_HITS if compact else 0) | (VG_TRACE_RAYS_PD if pd else 0)
        self._chk(self.L.vg_trace_batch_device(self.h, C.c_void_p(d_rays_ptr), C.c_int64(n), C.c_void_p(d_hits_ptr), C.c_uint32(flags)))

    def render(self, iter_begin: int, iter_end: int, fetch: bool = True, out: np.ndarray | None = None):
        """vg_render. `out`: a caller-owned (yres, xres, 3) float32 buffer for the frame; if it is page-locked (e.g. a
        torch pin_memory tensor's numpy view) the library DMAs straight into it."""
        if out is not None:
            assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (self.yres, self.xres, 3)
            fb = out
        else:
            fb = np.zeros((self.yres, self.xres, 3), np.float32) if fetch else None
        self._chk(self.L.vg_render(self.h, iter_begin, iter_end, _p(fb)))
        return fb

    def clear(self):
        self._chk(self.L.vg_clear_framebuffer(self.h))

    def framebuffer_ptr(self) -> int:
        p = C.c_void_p()
        self._chk(self.L.vg_framebuffer_device(self.h, C.byref(p)))
        return p.value

    def build_qbvh(self, boxes: np.ndarray, centroids: np.ndarray, leaf_max: int = 16):
        """vg_build_qbvh: (n,6) boxes {min,max}, (n,3) centroids -> (nodes NODE_DTYPE[], idx int32[n], bounds float32[6])."""
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        centroids = np.ascontiguousarray(centroids, np.float32).reshape(-1, 3)
        n = len(boxes)
        idx = np.zeros(n, np.int32)
        b6 = np.zeros(6, np.float32)
        nn = C.c_int()
        self._chk(self.L.vg_build_qbvh(self.h, _p(boxes), _p(centroids), n, leaf_max, _p(idx), _p(b6), C.byref(nn)))
        nodes = np.zeros(nn.value, NODE_DTYPE)
        self._chk(self.L.vg_build_qbvh_nodes(self.h, _p(nodes), nn.value))
        return nodes, idx, b6

    # -- texture store (vg_texture_*) ---------------------------------------------------------------
    def texture_upload(self, pixels_bottom_up: np.ndarray) -> int:
        px = np.ascontiguousarray(pixels_bottom_up, np.uint8)
        assert px.ndim == 3 and px.shape[2] == 3
        tid = C.c_int(-1)
        self._chk(self.L.vg_texture_upload(self.h, _p(px), int(px.shape[1]), int(px.shape[0]), C.byref(tid)))
        return tid.value

    def texture_levels(self, tex_id: int):
        """The pyramid the device built: list of (h, w, 3) uint8 arrays, rows bottom-up."""
        n = C.c_int()
        self._chk(self.L.vg_texture_levels(self.h, tex_id, C.byref(n)))
        out = []
        for l in range(n.value):
            w, h = C.c_int(), C.c_int()
            self._chk(self.L.vg_texture_read_level(self.h, tex_id, l, C.byref(w), C.byref(h), None))
            a = np.zeros((h.value, w.value, 3), np.uint8)
            self._chk(self.L.vg_texture_read_level(self.h, tex_id, l, C.byref(w), C.byref(h), _p(a)))
            out.append(a)
        return out

    def texture_sample(self, tex_id: int, coords: np.ndarray, trilinear: bool = False) -> np.ndarray:
        """vg_texture_sample_batch: coords (n, 8) {U, V, Dduvdx[2], Dduvdy[2], PixelDelta[2]} -> (n, 3)."""
        coords = np.ascontiguousarray(coords, np.float32).reshape(-1, 8)
        out = np.zeros((len(coords), 3), np.float32)
        self._chk(self.L.vg_texture_sample_batch(self.h, tex_id, 1 if trilinear else 0, _p(coords), C.c_int64(len(coords)), _p(out)))
        return out

    # -- multi-GPU (vg_comm_*, vg_gather_frame) -----------------------------------------------------
    def comm_unique_id(self) -> bytes:
        """Rank 0: the NCCL id every rank passes to comm_init (the caller distributes it)."""
        _torch_nccl_first()
        buf = C.create_string_buffer(VG_COMM_ID_BYTES)
        self._chk(self.L.vg_comm_unique_id(self.h, buf))
        return buf.raw

    def comm_init(self, rank: int, world: int, uid: bytes | None):
        """Collective: NCCL communicator on this context's GPU + the image partition of set_partition(rank, world)."""
        if world > 1:
            _torch_nccl_first()
        buf = C.create_string_buffer(uid, VG_COMM_ID_BYTES) if uid is not None else None
        self._chk(self.L.vg_comm_init(self.h, rank, world, buf))

    def gather_frame(self, out: np.ndarray | None = None):
        """Collective: owned pixels -> rank 0 (NCCL send/recv + the library's own pack/scatter kernels); rank 0 receives the
        complete frame in `out` (yres, xres, 3) float32."""
        if out is not None:
            assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (self.yres, self.xres, 3)
        self._chk(self.L.vg_gather_frame(self.h, _p(out)))
        return out

    def render_frame(self, table: np.ndarray, iter_begin: int, iter_end: int, out: np.ndarray | None = None, clear: bool = True):
        """vg_render_frame: scramble upload + (clear) + render + (multi-GPU gather) + frame download as one pipelined call."""
        table = np.ascontiguousarray(table, np.uint64)
        if out is not None:
            assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (self.yres, self.xres, 3)
        self._chk(self.L.vg_render_frame(self.h, _p(table), C.c_int64(table.shape[0]), iter_begin, iter_end, 1 if clear else 0, _p(out)))
        return out

    def measure_peaks(self) -> dict:
        """Streaming-read bandwidth of HBM, L2 and L1 on this GPU (csrc/peaks.cu)."""
        pk = VgPeaks()
        self._chk(self.L.vg_measure_peaks(self.h, C.byref(pk)))
        return {k: getattr(pk, k) for k, _ in VgPeaks._fields_ if k != "pad_"}

    def stats(self) -> dict:
        s = VgStats()
        self._chk(self.L.vg_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in VgStats._fields_}

    def reset_stats(self):
        self._chk(self.L.vg_reset_stats(self.h))


def _torch_nccl_first():
    """The library binds NCCL by soname (dlopen "libnccl.so.2") when the first communicator is made. PyTorch ships its own,
    newer copy under the same soname; the dynamic linker keeps whichever was loaded first for both. If this process is going to
    import torch at all, it has to do so BEFORE the library binds the system copy, or libtorch_cuda fails to resolve its symbols."""
    import importlib.util
    import sys
    if "torch" not in sys.modules and importlib.util.find_spec("torch") is not None:
        import torch  # noqa: F401


def device_count() -> int:
    return load_library().vg_device_count()


def owned_pixels(xres: int, yres: int, rank: int, world: int, pixel_block: bool = True) -> np.ndarray:
    """vg_owned_pixels: the library's own tile-major pixel list of `rank` (host code, no GPU needed)."""
    L = load_library()
    L.vg_owned_pixels.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
    n = L.vg_owned_pixels(xres, yres, rank, world, 1 if pixel_block else 0, None, 0)
    if n < 0:
        raise ValueError("vg_owned_pixels: bad arguments")
    out = np.zeros(n, np.int32)
    L.vg_owned_pixels(xres, yres, rank, world, 1 if pixel_block else 0, _p(out), n)
    return out
